"""CPU tests: the oracle (oracle/match_oracle.c) against (a) the committed golden vectors produced by the reference's
own nanoflann engine, (b) that engine live when oracle/_ref/ is present, (c) an independent numpy brute force,
(d) cv2.BFMatcher."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from metricsfm_b200 import synth

RATIOS = {"r50": 0.5, "r60": 0.6, "r85": 0.85}


def numpy_knn2(ref, qry):
    r = ref.astype(np.int64)
    q = qry.astype(np.int64)
    d = (q * q).sum(1)[:, None] + (r * r).sum(1)[None, :] - 2 * q @ r.T
    order = np.argsort(d, axis=1, kind="stable")[:, :2]  # stable => lowest index on ties
    return order.astype(np.int32), np.take_along_axis(d, order, 1).astype(np.float32), d


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference_golden(oracle_mod, golden, case):
    g = golden[case]
    ids, dists = oracle_mod.knn2_u8(g["ref"], g["qry"])
    # distances: identical to both builds of the reference engine
    np.testing.assert_array_equal(dists, g["fm_dists"])
    np.testing.assert_array_equal(dists, g["plain_dists"])
    # ids: identical to the lowest-index (NANOFLANN_FIRST_MATCH) build everywhere ...
    np.testing.assert_array_equal(ids, g["fm_ids"])
    # ... and to the stock build wherever the two nearest distances are not tied
    untied = (g["plain_dists"][:, 0] != g["plain_dists"][:, 1])
    np.testing.assert_array_equal(ids[untied, 0], g["plain_ids"][untied, 0])
    for tag, th in RATIOS.items():
        pairs, _ = oracle_mod.ratio_select(ids, dists, g["ref"].shape[0], th, orientation=0)
        exp = g["fm_pairs_" + tag]
        if exp.shape[0] == 1 and exp[0, 0] == -7:      # reference returned false (<20 keypoints)
            assert pairs is None
        else:
            np.testing.assert_array_equal(pairs, exp)
            np.testing.assert_array_equal(pairs, g["plain_pairs_" + tag])


def test_oracle_vs_reference_nanoflann_live(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    col = synth.Collection(1500, seed=21)
    a, b = col.image_u8(0, 1500), col.image_u8(1, 1203)
    ids, dists = oracle_mod.knn2_u8(a, b)
    rids, rdists = oracle_mod.ref_knn2(a.astype(np.float32), b.astype(np.float32))
    np.testing.assert_array_equal(dists, rdists)
    untied = rdists[:, 0] != rdists[:, 1]
    np.testing.assert_array_equal(ids[untied], rids[untied])
    exp = oracle_mod.ref_match(a.astype(np.float32), b.astype(np.float32), 0.5, 20)
    got = oracle_mod.match_pair_u8(a, b, 0.5)["pairs"]
    np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("m,n", [(1, 5), (2, 3), (19, 20), (20, 20), (127, 129), (300, 64)])
def test_oracle_vs_numpy(oracle_mod, m, n):
    rng = np.random.default_rng(m * 1000 + n)
    ref = rng.integers(0, 256, size=(m, 128), dtype=np.uint8)
    qry = rng.integers(0, 256, size=(n, 128), dtype=np.uint8)
    if m > 4:
        ref[3] = ref[1]
    ids, dists = oracle_mod.knn2_u8(ref, qry)
    nid, nd, full = numpy_knn2(ref, qry)
    if m >= 2:
        np.testing.assert_array_equal(ids, nid)
        np.testing.assert_array_equal(dists, nd)
    else:
        np.testing.assert_array_equal(ids[:, 0], nid[:, 0])
        assert (ids[:, 1] == -1).all() and np.isinf(dists[:, 1]).all()
    cb, cd = oracle_mod.colbest_u8(ref, qry)
    np.testing.assert_array_equal(cb, np.argmin(full, axis=0).astype(np.int32))
    np.testing.assert_array_equal(cd, full.min(axis=0).astype(np.float32))


def test_oracle_f32_equals_u8_on_integer_rows(oracle_mod):
    col = synth.Collection(400, seed=5)
    a, b = col.image_u8(0), col.image_u8(1)
    i8, d8 = oracle_mod.knn2_u8(a, b)
    i32, d32 = oracle_mod.knn2_f32(a.astype(np.float32), b.astype(np.float32))
    np.testing.assert_array_equal(i8, i32)
    np.testing.assert_array_equal(d8, d32)


def test_oracle_vs_cv2_bfmatcher(oracle_mod):
    cv2 = pytest.importorskip("cv2")
    col = synth.Collection(500, seed=6)
    a, b = col.image_u8(0), col.image_u8(1)
    ids, dists = oracle_mod.knn2_u8(a, b)
    bf = cv2.BFMatcher(cv2.NORM_L2)
    res = bf.knnMatch(b.astype(np.float32), a.astype(np.float32), k=2)
    cid = np.array([[m[0].trainIdx, m[1].trainIdx] for m in res], dtype=np.int32)
    cd = np.rint(np.array([[m[0].distance, m[1].distance] for m in res], dtype=np.float64) ** 2).astype(np.float32)
    np.testing.assert_array_equal(dists, cd)
    untied = dists[:, 0] != dists[:, 1]
    np.testing.assert_array_equal(ids[untied, 0], cid[untied, 0])


def test_ratio_rules(oracle_mod):
    ids = np.array([[3, 4], [5, 6], [7, 8], [9, -1], [1, 2]], dtype=np.int32)
    dists = np.array([[0, 0], [0, 10], [6, 10], [1, np.inf], [5, 10]], dtype=np.float32)
    pairs, good = oracle_mod.ratio_select(ids, dists, 30, 0.6, min_keypoints=0, ratio_good=0.5)
    # 0/0 NaN rejected; 0/10 accepted; 0.6 !< 0.6 rejected (strict); missing second neighbour rejected; 0.5 accepted
    np.testing.assert_array_equal(pairs, [[5, 1], [1, 4]])
    np.testing.assert_array_equal(good, [1, 0])  # ratio_good is strict as well: 0.5 !< 0.5
    pairs, _ = oracle_mod.ratio_select(ids, dists, 30, 0.6, min_keypoints=0, orientation=1)
    np.testing.assert_array_equal(pairs, [[1, 5], [4, 1]])
    assert oracle_mod.ratio_select(ids, dists, 30, 0.6, min_keypoints=20)[0] is None  # N=5 < 20
    cb = np.full((30,), -1, np.int32)
    cb[5] = 1
    pairs, _ = oracle_mod.ratio_select(ids, dists, 30, 0.6, min_keypoints=0, col_best=cb)
    np.testing.assert_array_equal(pairs, [[5, 1]])
    pairs, _ = oracle_mod.ratio_select(ids, dists, 30, 0.6, min_keypoints=0, max_dist_sq=5.0)
    np.testing.assert_array_equal(pairs, [[5, 1]])


def test_quantizer_golden(oracle_mod, golden):
    g = golden["float_unit"]
    q = oracle_mod.quantize_f32(g["unit"], 512.0)
    np.testing.assert_array_equal(q, g["q512"])
    np.testing.assert_array_equal(q, np.clip(np.rint(g["unit"] * np.float32(512.0)), 0, 255).astype(np.uint8))
    edge = np.zeros((1, 128), np.float32)
    edge[0, :6] = [-3.0, np.nan, 0.49, 0.5, 254.5, 1e9]
    np.testing.assert_array_equal(oracle_mod.quantize_f32(edge, 1.0)[0, :6], [0, 0, 0, 0, 254, 255])


def test_mutual_subset(oracle_mod):
    col = synth.Collection(600, seed=9)
    a, b = col.image_u8(0), col.image_u8(1)
    one = oracle_mod.match_pair_u8(a, b, 0.85)["pairs"]
    mut = oracle_mod.match_pair_u8(a, b, 0.85, mutual=True)["pairs"]
    s1 = {tuple(p) for p in one}
    assert {tuple(p) for p in mut} <= s1
    assert len(np.unique(mut[:, 0])) == len(mut)  # a reference row is claimed by at most one query
