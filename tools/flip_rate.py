"""Float-regime report (north_star: "where the reference uses float-normalized descriptors, the ratio-boundary flip rate
is reported"): match lists of the u8 path (q = min(255, rint(512 x)), what the packer stores — the CUDA path is
bit-identical to the u8 oracle by the GPU parity tests) versus exact fp32 brute force on the unit-norm float rows
(the reference's CUDASIFT container, feature_extractor_cuda_sift.cpp:75-80).  CPU only; prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metricsfm_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main(rows=8192, n_pairs=3):
    col = synth.Collection(rows, seed=0)
    out = {}
    for ratio in (0.5, 0.6, 0.85):
        flips = total = nn0_diff = nn1_diff = nrows = 0
        for p in range(n_pairs):
            a, b = col.image_unit(2 * p), col.image_unit(2 * p + 1)
            fi, fd = oracle.knn2_f32(a, b)
            qa, qb = oracle.quantize_f32(a, 512.0), oracle.quantize_f32(b, 512.0)
            qi, qd = oracle.knn2_u8(qa, qb)
            fm, _ = oracle.ratio_select(fi, fd, rows, ratio)
            qm, _ = oracle.ratio_select(qi, qd, rows, ratio)
            sf, sq = {tuple(x) for x in fm}, {tuple(x) for x in qm}
            flips += len(sf ^ sq)
            total += len(sf)
            nn0_diff += int((fi[:, 0] != qi[:, 0]).sum())
            nn1_diff += int((fi[:, 1] != qi[:, 1]).sum())
            nrows += rows
        out[f"ratio_{ratio}"] = {"matches_fp32": total, "symmetric_difference": flips, "flip_rate": flips / max(total, 1)}
    out["nn0_identity_diff_rate"] = nn0_diff / nrows
    out["nn1_identity_diff_rate"] = nn1_diff / nrows
    out["rows"] = rows
    out["pairs"] = n_pairs
    out["tolerance_target"] = 1e-4
    print(json.dumps(out))


if __name__ == "__main__":
    main()
