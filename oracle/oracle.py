"""ctypes front-end of the CPU oracle (oracle/match_oracle.c) and of the compiled reference engine
(oracle/_ref/libref_nanoflann*.so, built from the reference's vendored nanoflann headers).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs.  Nothing under metricsfm_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
_REF_FM_SO = os.path.join(_HERE, "_ref", "libref_nanoflann_firstmatch.so")

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference engine (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/SfM/src/utils"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build(ref=False)   # make: a no-op when _build/liboracle.so is newer than match_oracle.c
        L = C.CDLL(_ORACLE_SO)
        L.oracle_max_threads.restype = C.c_int
        L.oracle_quantize_f32.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_float, _u8p]
        L.oracle_knn2_u8.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, _i32p, _f32p]
        L.oracle_knn2_f32.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32, _i32p, _f32p]
        L.oracle_colbest_u8.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, _i32p, _f32p]
        L.oracle_colbest_f32.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32, _i32p, _f32p]
        L.oracle_ratio_select.argtypes = [_i32p, _f32p, C.c_int32, C.c_int32, C.c_float, C.c_float, _i32p, C.c_int32,
                                          C.c_int32, C.c_float, _i32p, _u8p]
        L.oracle_ratio_select.restype = C.c_int32
        L.oracle_ratio_select_rule.argtypes = [_i32p, _f32p, C.c_int32, C.c_int32, C.c_float, C.c_float, _i32p, C.c_int32,
                                               C.c_int32, C.c_float, C.c_int32, _i32p, _u8p]
        L.oracle_ratio_select_rule.restype = C.c_int32
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def use_all_host_threads() -> int:
    """Size the OpenMP teams of the oracle and of the reference engine from the CPUs this process may run on (launchers
    like torchrun export OMP_NUM_THREADS=1).  Returns the thread count."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_threads(n)
    if ref_available():
        ref_lib().ref_nanoflann_set_threads(n)
    return n


def quantize_f32(desc: np.ndarray, scale: float) -> np.ndarray:
    desc = np.ascontiguousarray(desc, dtype=np.float32)
    assert desc.ndim == 2 and desc.shape[1] == 128
    out = np.empty(desc.shape, dtype=np.uint8)
    lib().oracle_quantize_f32(_p(desc, _f32p), desc.shape[0], 128, scale, _p(out, _u8p))
    return out


def knn2_u8(ref: np.ndarray, query: np.ndarray):
    """FLANN-layout exact 2-NN: returns ids [N,2] int32, dists [N,2] float32 (squared L2)."""
    ref = np.ascontiguousarray(ref, dtype=np.uint8).reshape(-1, 128)
    query = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 128)
    n = query.shape[0]
    ids = np.empty((n, 2), dtype=np.int32)
    dists = np.empty((n, 2), dtype=np.float32)
    lib().oracle_knn2_u8(_p(ref, _u8p), ref.shape[0], _p(query, _u8p), n, _p(ids, _i32p), _p(dists, _f32p))
    return ids, dists


def knn2_f32(ref: np.ndarray, query: np.ndarray):
    ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1, 128)
    query = np.ascontiguousarray(query, dtype=np.float32).reshape(-1, 128)
    n = query.shape[0]
    ids = np.empty((n, 2), dtype=np.int32)
    dists = np.empty((n, 2), dtype=np.float32)
    lib().oracle_knn2_f32(_p(ref, _f32p), ref.shape[0], _p(query, _f32p), n, _p(ids, _i32p), _p(dists, _f32p))
    return ids, dists


def colbest_u8(ref: np.ndarray, query: np.ndarray):
    ref = np.ascontiguousarray(ref, dtype=np.uint8).reshape(-1, 128)
    query = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 128)
    m = ref.shape[0]
    best = np.empty((m,), dtype=np.int32)
    dist = np.empty((m,), dtype=np.float32)
    lib().oracle_colbest_u8(_p(ref, _u8p), m, _p(query, _u8p), query.shape[0], _p(best, _i32p), _p(dist, _f32p))
    return best, dist


def colbest_f32(ref: np.ndarray, query: np.ndarray):
    ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1, 128)
    query = np.ascontiguousarray(query, dtype=np.float32).reshape(-1, 128)
    m = ref.shape[0]
    best = np.empty((m,), dtype=np.int32)
    dist = np.empty((m,), dtype=np.float32)
    lib().oracle_colbest_f32(_p(ref, _f32p), m, _p(query, _f32p), query.shape[0], _p(best, _i32p), _p(dist, _f32p))
    return best, dist


def ratio_select(ids, dists, m_ref: int, ratio: float, *, max_dist_sq: float = 0.0, col_best=None, min_keypoints: int = 20,
                 orientation: int = 0, ratio_good: float = 0.0, reject_gt: bool = False):
    """Returns (pairs [n,2] int32, good_flags [n] uint8) or (None, None) when the <min_keypoints gate rejects.
    reject_gt: SLAMGPS::FeatureMatching's rule (slam_gps.cc:470-477), accept iff !(d0/d1 > ratio)."""
    ids = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1, 2)
    dists = np.ascontiguousarray(dists, dtype=np.float32).reshape(-1, 2)
    n = ids.shape[0]
    pairs = np.empty((max(n, 1), 2), dtype=np.int32)
    flags = np.zeros((max(n, 1),), dtype=np.uint8)
    cb = None
    if col_best is not None:
        cb = np.ascontiguousarray(col_best, dtype=np.int32)
    cnt = lib().oracle_ratio_select_rule(_p(ids, _i32p), _p(dists, _f32p), m_ref, n, ratio, max_dist_sq,
                                         _p(cb, _i32p) if cb is not None else None, min_keypoints, orientation, ratio_good,
                                         1 if reject_gt else 0, _p(pairs, _i32p), _p(flags, _u8p))
    if cnt < 0:
        return None, None
    return pairs[:cnt].copy(), flags[:cnt].copy()


def match_pair_u8(ref, query, ratio: float, *, mutual: bool = False, max_dist_sq: float = 0.0, min_keypoints: int = 20,
                  orientation: int = 0, ratio_good: float = 0.0, reject_gt: bool = False):
    """Whole path for one pair.  Returns dict(ok, ids, dists, pairs, good)."""
    ref = np.ascontiguousarray(ref, dtype=np.uint8).reshape(-1, 128)
    query = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 128)
    if ref.shape[0] < min_keypoints or query.shape[0] < min_keypoints:
        return dict(ok=False, ids=None, dists=None, pairs=np.empty((0, 2), np.int32), good=np.empty((0,), np.uint8))
    ids, dists = knn2_u8(ref, query)
    cb = colbest_u8(ref, query)[0] if mutual else None
    pairs, good = ratio_select(ids, dists, ref.shape[0], ratio, max_dist_sq=max_dist_sq, col_best=cb,
                               min_keypoints=min_keypoints, orientation=orientation, ratio_good=ratio_good, reject_gt=reject_gt)
    return dict(ok=True, ids=ids, dists=dists, pairs=pairs, good=good)


# ----------------------------------------------------------------------------------------------------------------
# The reference's own engine (nanoflann exact KD-tree), compiled from /root/reference by `make ref`.
# ----------------------------------------------------------------------------------------------------------------
_ref_libs = {}


def ref_available(first_match: bool = False) -> bool:
    return os.path.exists(_REF_FM_SO if first_match else _REF_SO)


def ref_lib(first_match: bool = False) -> C.CDLL:
    key = bool(first_match)
    if key not in _ref_libs:
        path = _REF_FM_SO if first_match else _REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(path)
        L.ref_nanoflann_max_threads.restype = C.c_int
        L.ref_nanoflann_build.argtypes = [_f32p, C.c_int32]
        L.ref_nanoflann_build.restype = C.c_void_p
        L.ref_nanoflann_free.argtypes = [C.c_void_p]
        L.ref_nanoflann_knn2.argtypes = [C.c_void_p, _f32p, C.c_int32, _i32p, _f32p]
        L.ref_nanoflann_match.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32, C.c_float, C.c_int32, _i32p]
        L.ref_nanoflann_match.restype = C.c_int32
        _ref_libs[key] = L
    return _ref_libs[key]


def ref_knn2(desc1_f32: np.ndarray, desc2_f32: np.ndarray, first_match: bool = False):
    """Reference engine: tree on image 1, 2-NN of every row of image 2 (feature_matching.cpp:319-337)."""
    L = ref_lib(first_match)
    d1 = np.ascontiguousarray(desc1_f32, dtype=np.float32).reshape(-1, 128)
    d2 = np.ascontiguousarray(desc2_f32, dtype=np.float32).reshape(-1, 128)
    ids = np.empty((d2.shape[0], 2), dtype=np.int32)
    dists = np.empty((d2.shape[0], 2), dtype=np.float32)
    h = L.ref_nanoflann_build(_p(d1, _f32p), d1.shape[0])
    try:
        L.ref_nanoflann_knn2(h, _p(d2, _f32p), d2.shape[0], _p(ids, _i32p), _p(dists, _f32p))
    finally:
        L.ref_nanoflann_free(h)
    return ids, dists


def ref_match(desc1_f32, desc2_f32, th_ratio: float = 0.5, th_reject: int = 20, first_match: bool = False):
    """Reference overload feature_matching.cpp:319-342 end to end: returns pairs [n,2] or None if gated."""
    L = ref_lib(first_match)
    d1 = np.ascontiguousarray(desc1_f32, dtype=np.float32).reshape(-1, 128)
    d2 = np.ascontiguousarray(desc2_f32, dtype=np.float32).reshape(-1, 128)
    pairs = np.empty((max(d2.shape[0], 1), 2), dtype=np.int32)
    n = L.ref_nanoflann_match(_p(d1, _f32p), d1.shape[0], _p(d2, _f32p), d2.shape[0], th_ratio, th_reject, _p(pairs, _i32p))
    if n < 0:
        return None
    return pairs[:n].copy()
