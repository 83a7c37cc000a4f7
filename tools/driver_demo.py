"""End-to-end run of the fine-matching-graph driver at BASELINE config #2 scale (100 images x 8192 SIFT-128, exhaustive
unordered pairs) through the reference's on-disk formats: <idx>_feature files in, <idx1>_match / match_index.txt /
graph_matching.txt out (msfm_build_match_graph = FineMatchingGraph::BuildMatchGraph, fine_matching_graph.cc:40-194).
Keypoints of images sharing scene-pool rows are given a common two-view-style geometry only implicitly (random
positions), so the GPU geo-verification rejects most pairs: this measures the pipeline, not a reconstruction.
Prints one JSON line.  GPU box only."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metricsfm_b200 import store, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--geo", type=int, default=1)
    ap.add_argument("--u8", type=int, default=0, help="1: store CV_8UC1 descriptors instead of the reference's CV_32FC1")
    args = ap.parse_args()
    fold = tempfile.mkdtemp(prefix="msfm_demo_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        col = synth.Collection(args.rows, seed=0)
        rng = np.random.default_rng(1)
        t0 = time.perf_counter()
        for i in range(args.images):
            d = col.image_u8(i)
            xy = rng.uniform(0, 4000, size=(args.rows, 2)).astype(np.float32)
            store.feature_write(store.feature_path(fold, i), rows=4000, cols=4000, xy_pixel=xy, desc=d if args.u8 else d.astype(np.float32))
        t_write = time.perf_counter() - t0
        adj = [[j for j in range(i + 1, args.images)] for i in range(args.images)]   # unordered pairs, grouped by idx1
        offs = np.cumsum([0] + [len(a) for a in adj]).astype(np.int64)
        lst = np.array([j for a in adj for j in a], np.int32)
        t0 = time.perf_counter()
        store.build_match_graph(fold, offs, lst, mutual=True, min_keypoints=20, geo_verify=bool(args.geo), min_good=0)
        t_graph = time.perf_counter() - t0
        g = store.graph_read(fold, args.images)
        n_pairs = int(len(lst))
        files = [f for f in os.listdir(fold) if f.endswith("_match")]
        out = {"images": args.images, "rows": args.rows, "pairs": n_pairs, "descriptor_type": "CV_8UC1" if args.u8 else "CV_32FC1",
               "feature_bytes": int(sum(os.path.getsize(store.feature_path(fold, i)) for i in range(args.images))),
               "write_features_s": t_write, "build_match_graph_s": t_graph, "pairs_per_s_end_to_end": n_pairs / t_graph,
               "geo_verify": bool(args.geo), "pairs_with_matches": int((g > 0).sum()), "matches_written": int(g.sum()),
               "match_files": len(files),
               "what": "msfm_build_match_graph: read feature files (tmpfs), stage in HBM, match all pairs (ratio 0.85/0.6 + mutual), "
                       "GPU geo-verification, write the reference's match files"}
        print(json.dumps(out))
    finally:
        shutil.rmtree(fold, ignore_errors=True)


if __name__ == "__main__":
    main()
