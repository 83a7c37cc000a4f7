"""Generate the committed golden vectors under tests/golden/ by running THE REFERENCE'S OWN exact kNN engine
(nanoflann, compiled from /root/reference by `make -C oracle ref`) on small seeded inputs.

Run in the build container (where /root/reference exists):   python tests/golden/gen_golden.py
The .npz files hold inputs and the reference outputs; tests/test_oracle.py pins the oracle to them on any box and
tests/test_gpu_parity.py pins the CUDA path to them on the GPU box (where /root/reference does not exist).

Reference path exercised: FeatureMatching::KNNMatchingWithGeoVerify(kp1, my_kd_tree_t*, kp2, descriptors2, matches)
kNN + ratio part, /root/reference/SfM/src/feature/feature_matching.cpp:319-350, with the ratio thresholds of the
callers (0.5 feature_matching.cpp:322; 0.6 / 0.85 fine_matching_graph.cc:42-43).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from metricsfm_b200 import synth  # noqa: E402


def ref_outputs(ref_u8, qry_u8, first_match):
    ids, dists = oracle.ref_knn2(ref_u8.astype(np.float32), qry_u8.astype(np.float32), first_match=first_match)
    out = {"ids": ids, "dists": dists}
    for name, th in (("r50", 0.5), ("r60", 0.6), ("r85", 0.85)):
        pairs = oracle.ref_match(ref_u8.astype(np.float32), qry_u8.astype(np.float32), th_ratio=th, th_reject=20,
                                 first_match=first_match)
        out["pairs_" + name] = pairs if pairs is not None else np.full((1, 2), -7, np.int32)  # -7 sentinel = gated
    return out


def case_basic():
    col = synth.Collection(300, seed=11)
    return col.image_u8(0, 300), col.image_u8(1, 257)


def case_ragged():
    col = synth.Collection(700, seed=12)
    return col.image_u8(3, 641), col.image_u8(4, 130)


def case_ties():
    rng = np.random.default_rng(1234)
    ref = rng.integers(0, 60, size=(96, 128), dtype=np.uint8)
    ref[10] = ref[5]            # duplicate reference rows -> equal distances, lowest index must win
    ref[77] = ref[5]
    ref[40] = 0                 # all-zero row
    ref[41] = 0
    ref[50] = 255               # saturated row
    qry = rng.integers(0, 60, size=(64, 128), dtype=np.uint8)
    qry[0] = ref[5]             # exact hit on a triplicated row: d0 = d1 = 0 -> ratio NaN -> rejected
    qry[1] = ref[20]            # exact hit on a unique row: d0 = 0 < ratio*d1 -> accepted
    qry[2] = 0
    qry[3] = 255
    qry[4] = qry[1]             # duplicate queries (matters for the mutual check)
    return ref, qry


def case_small_gate():
    col = synth.Collection(64, seed=13)
    return col.image_u8(0, 19), col.image_u8(1, 40)   # 19 < th_reject=20 -> reference returns false


def case_min_ok():
    col = synth.Collection(64, seed=14)
    return col.image_u8(0, 20), col.image_u8(1, 21)


def main():
    if not oracle.ref_available(True):
        oracle.build(ref=True)
    cases = {"basic": case_basic, "ragged": case_ragged, "ties": case_ties, "small_gate": case_small_gate, "min_ok": case_min_ok}
    for name, fn in cases.items():
        ref, qry = fn()
        fm = ref_outputs(ref, qry, first_match=True)     # header compiled with its NANOFLANN_FIRST_MATCH switch
        plain = ref_outputs(ref, qry, first_match=False)  # header as the reference builds it (ties: visit order)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), ref=ref, qry=qry,
                            **{"fm_" + k: v for k, v in fm.items()}, **{"plain_" + k: v for k, v in plain.items()})
        print(name, ref.shape, qry.shape, {k: v.shape for k, v in fm.items()})
    # float regime: unit-norm rows (CUDASIFT regime) and the quantiser's expected output
    col = synth.Collection(200, seed=15)
    unit = col.image_unit(0, 200)
    np.savez_compressed(os.path.join(HERE, "float_unit.npz"), unit=unit, q512=oracle.quantize_f32(unit, 512.0))
    print("float_unit", unit.shape)


if __name__ == "__main__":
    main()
