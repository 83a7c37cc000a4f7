"""Float-regime report (north_star: "where the reference uses float-normalized descriptors, the ratio-boundary flip rate
is reported and must stay within a stated tolerance (<= 1e-4 of matches)").

Inputs: unit-norm float rows (the reference's CUDASIFT container, feature_extractor_cuda_sift.cpp:75-80), uploaded with
scale 512.  Reference answer: exact fp32 brute force (oracle.knn2_f32: squared L2 accumulated in index order,
nanoflann.hpp:376-383) + ratio test (+ mutual).  Compared:
  quantised  the u8 path alone (q = min(255, rint(512 x)), what the packer stores)
  rescored   the CUDA path with msfm_config.keep_float and msfm_params.rescore_band (rows near a ratio threshold are
             decided on exact fp32 distances)
Run on a GPU box (`python tools/flip_rate.py [--rows 8192] [--pairs 6] [--band 0.03]`); prints one JSON line.  With
--cpu-only the rescored column is skipped (no GPU needed)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metricsfm_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


_cache = {}


def lists_fp32(key, a, b, ratio, good, mutual):
    if key not in _cache:  # the fp32 brute force is the slow part: once per pair
        _cache[key] = (oracle.knn2_f32(a, b), oracle.colbest_f32(a, b)[0])
    (ids, dists), cb = _cache[key]
    return oracle.ratio_select(ids, dists, a.shape[0], ratio, col_best=cb if mutual else None, ratio_good=good)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--pairs", type=int, default=6)
    ap.add_argument("--band", type=float, default=0.03)
    ap.add_argument("--cpu-only", action="store_true")
    args = ap.parse_args()
    col = synth.Collection(args.rows, seed=0)
    n_img = 2 * args.pairs
    imgs = [col.image_unit(i) for i in range(n_img)]
    pairs = [(2 * p, 2 * p + 1) for p in range(args.pairs)]
    out = {"rows": args.rows, "pairs": args.pairs, "band": args.band, "tolerance_target": 1e-4, "scale": 512.0}
    m = None
    if not args.cpu_only:
        from metricsfm_b200.matcher import Matcher
        m = Matcher(device=0, max_images=n_img, arena_rows=n_img * (args.rows + 256), keep_float=True)
        for i, x in enumerate(imgs):
            m.upload(i, x, scale=512.0)
    for mutual in (False, True):
        for ratio in (0.5, 0.6, 0.85):
            total = fl_q = fl_r = 0
            res_q = m.match_pairs(pairs, ratio, mutual=mutual) if m else None
            ms_q = m.timing()["total_ms"] if m else None
            res_r = m.match_pairs(pairs, ratio, mutual=mutual, rescore_band=args.band) if m else None
            ms_r = m.timing()["total_ms"] if m else None
            for p, (r, q) in enumerate(pairs):
                ref = {tuple(x) for x in lists_fp32((r, q), imgs[r], imgs[q], ratio, 0.0, mutual)}
                total += len(ref)
                if m:
                    fl_q += len(ref ^ {tuple(x) for x in res_q.pair(p)})
                    fl_r += len(ref ^ {tuple(x) for x in res_r.pair(p)})
                else:
                    qa, qb = oracle.quantize_f32(imgs[r], 512.0), oracle.quantize_f32(imgs[q], 512.0)
                    fl_q += len(ref ^ {tuple(x) for x in oracle.match_pair_u8(qa, qb, ratio, mutual=mutual)["pairs"]})
            key = f"ratio_{ratio}{'_mutual' if mutual else ''}"
            out[key] = {"matches_fp32": total, "flips_quantised": fl_q, "flip_rate_quantised": fl_q / max(total, 1)}
            if m:
                out[key].update({"flips_rescored": fl_r, "flip_rate_rescored": fl_r / max(total, 1),
                                 "device_ms_quantised": ms_q, "device_ms_rescored": ms_r})
    if m:
        m.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
