"""GPU tests of the single-process multi-GPU pair scheduler (include/msfm_multi.h): staging with NCCL replication,
cost-balanced sharding, per-device matching threads, stitched results.  The 1-device cases run on any B200 box; the
N-device cases need `gpurun --gpus N` and are skipped otherwise."""
import numpy as np
import pytest

from metricsfm_b200 import synth

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def _collection(n, rows_max, seed):
    col = synth.Collection(rows_max, seed=seed)
    rows = [rows_max - (53 * i) % 400 for i in range(n)]
    rows[3] = 12          # gated (< 20 keypoints)
    rows[5] = 0           # empty
    return [np.ascontiguousarray(col.image_u8(i, r)) for i, r in enumerate(rows)]


def _cross_block_pairs(n):
    """Every image paired with a spread of partners, so that a shard reads images staged by every device."""
    pairs = [(i, j) for i in range(n) for j in range(n) if i != j and (i + 2 * j) % 3 == 0]
    return np.array(pairs, np.int32)


def _reference_lists(native_lib, imgs, pairs, **kw):
    from metricsfm_b200.matcher import Matcher
    total = sum(x.shape[0] for x in imgs) + 256 * len(imgs)
    with Matcher(device=0, max_images=len(imgs), arena_rows=total) as m:
        for i, x in enumerate(imgs):
            m.upload(i, x)
        return m.match_pairs(pairs, 0.85, **kw)


@pytest.mark.parametrize("ndev", [1, 2, 4, 8])
def test_multi_gpu_lists_equal_single_gpu_lists(oracle_mod, native_lib, ndev):
    """N-GPU match lists are bit-identical to the 1-GPU lists (and to the oracle on sampled pairs); every device's replica
    of the table — packed locally or received over NCCL — holds the host rows byte for byte."""
    if _device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import torch
    from metricsfm_b200.multi import MultiMatcher
    n = 24
    imgs = _collection(n, 1536, seed=29)
    pairs = _cross_block_pairs(n)
    kw = dict(ratio_good=0.6, mutual=True, min_keypoints=20)
    ref = _reference_lists(native_lib, imgs, pairs, **kw)
    cap = int(sum(imgs[q].shape[0] for _, q in pairs))
    pinned = [torch.from_numpy(x).pin_memory() if x.shape[0] else torch.from_numpy(x) for x in imgs]
    with MultiMatcher(list(range(ndev)), max_images=n, arena_rows=sum(x.shape[0] for x in imgs) + 256 * n) as mm:
        # two staging groups: the second one is in flight while the first group's pairs are matched
        g0, g1 = list(range(0, 10)), list(range(10, n))
        mm.upload_u8(g0, [pinned[i] for i in g0])
        mm.upload_u8(g1, [pinned[i] for i in g1])
        early = np.array([p for p in pairs if max(p) < 10], np.int32)
        r_early = mm.match_pairs(early, 0.85, capacity=cap, **kw)
        got = mm.match_pairs(pairs, 0.85, capacity=cap, **kw)
        t, per = mm.timing()
        assert t["n_devices"] == ndev and sum(p["match_launches"] > 0 for p in per) == ndev
        for d in range(ndev):
            for i in (0, 7, 11, n - 1):
                desc, _ = mm.download_packed(d, i)
                np.testing.assert_array_equal(desc, imgs[i], err_msg=f"device {d} image {i}")
    np.testing.assert_array_equal(got.offsets, ref.offsets)
    np.testing.assert_array_equal(got.ok, ref.ok)
    np.testing.assert_array_equal(got.matches, ref.matches)
    np.testing.assert_array_equal(got.good, ref.good)
    k = 0
    for p, pr in enumerate(pairs):
        if max(pr) < 10:
            np.testing.assert_array_equal(r_early.pair(k), ref.pair(p))
            k += 1
    for p in range(0, len(pairs), 17):
        r, q = pairs[p]
        exp = oracle_mod.match_pair_u8(imgs[r], imgs[q], 0.85, mutual=True, ratio_good=0.6)
        assert bool(got.ok[p]) == exp["ok"]
        np.testing.assert_array_equal(got.pair(p), exp["pairs"])


def test_multi_f32_upload_and_errors(native_lib):
    from metricsfm_b200.matcher import MsfmError
    from metricsfm_b200.multi import MultiMatcher
    ndev = min(_device_count(), 2)
    col = synth.Collection(900, seed=31)
    imgs = [np.ascontiguousarray(col.image_u8(i, 900 - 10 * i)) for i in range(6)]
    pairs = synth.exhaustive_pairs(6)
    ref = _reference_lists(native_lib, imgs, pairs, ratio_good=0.6, mutual=False)
    with MultiMatcher(list(range(ndev)), max_images=8, arena_rows=8 * 1024) as mm:
        mm.upload_f32(list(range(6)), [x.astype(np.float32) for x in imgs], scale=1.0)   # CV_32FC1 rows, integer-valued
        got = mm.match_pairs(pairs, 0.85, ratio_good=0.6, capacity=6000 * 15)
        np.testing.assert_array_equal(got.matches, ref.matches)
        np.testing.assert_array_equal(got.good, ref.good)
        with pytest.raises(MsfmError):
            mm.upload_u8([2], [imgs[2]])                         # already staged
        with pytest.raises(MsfmError):
            mm.match_pairs([(0, 7)], 0.85, capacity=1000)        # image 7 never staged
        with pytest.raises(MsfmError):
            mm.match_pairs(pairs, 0.85, capacity=10)             # result buffer too small
        mm.release_all()
        mm.upload_u8([0, 1], [imgs[0], imgs[1]])
        again = mm.match_pairs([(0, 1)], 0.85, ratio_good=0.6, capacity=2000)
        np.testing.assert_array_equal(again.pair(0), ref.pair(0))
