"""ctypes mirror of include/msfm_store.h and include/msfm_graph.h: the reference's on-disk formats around the matching
hot path (feature files, match files, resume index, graph files, pair lists) and the fine-matching-graph driver.
Thin marshalling only; the work happens in libmsfm_store.so / libmsfm_graph.so."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import GRAPH_LIB, STORE_LIB

STORE_SYMBOLS = [
    "msfm_feature_path", "msfm_feature_stat", "msfm_feature_read", "msfm_feature_write", "msfm_match_path", "msfm_match_append",
    "msfm_match_read", "msfm_match_index_missing", "msfm_match_index_append", "msfm_graph_write", "msfm_graph_read",
    "msfm_graph_recover", "msfm_pairs_all", "msfm_pairs_priori_xy", "msfm_init_graph_write", "msfm_init_graph_read",
    "msfm_similarity_invfile", "msfm_pairs_similarity_topk", "msfm_word_matches",
]
GRAPH_SYMBOLS = ["msfm_build_match_graph"]
ERR_CAPACITY = -4


class FeatureInfo(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("zoom_ratio", C.c_float), ("f_mm", C.c_float), ("f_pixel", C.c_float),
                ("gps_latitude", C.c_float), ("gps_longitude", C.c_float), ("maker_len", C.c_int32), ("model_len", C.c_int32),
                ("num_pts", C.c_int32), ("desc_rows", C.c_int32), ("desc_cols", C.c_int32), ("desc_type", C.c_int32),
                ("desc_elem_size", C.c_int32), ("keypoints_offset", C.c_int64), ("desc_offset", C.c_int64)]


class GraphOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("th_good", C.c_float), ("th_all", C.c_float), ("mutual", C.c_int32),
                ("min_keypoints", C.c_int32), ("min_good", C.c_int32), ("descriptor_scale", C.c_float), ("rescore_band", C.c_float),
                ("geo_verify", C.c_int32), ("geo_seed", C.c_uint32), ("max_batch_rows", C.c_int64), ("n_devices", C.c_int32),
                ("reserved", C.c_int32)]


VERIFY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_float), C.c_int32,
                        C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32))

_store = None
_graph = None


class StoreError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _store
    if _store is None:
        if not os.path.exists(STORE_LIB):
            raise RuntimeError(f"{STORE_LIB} not built (python -m metricsfm_b200.build)")
        _store = C.CDLL(STORE_LIB)
        for n in STORE_SYMBOLS:
            getattr(_store, n).restype = C.c_int
    return _store


def graph_lib() -> C.CDLL:
    global _graph
    if _graph is None:
        if not os.path.exists(GRAPH_LIB):
            raise RuntimeError(f"{GRAPH_LIB} not built (python -m metricsfm_b200.build)")
        _graph = C.CDLL(GRAPH_LIB)
        _graph.msfm_build_match_graph.restype = C.c_int
    return _graph


def _chk(rc: int, what: str):
    if rc != 0:
        raise StoreError(f"{what} failed with status {rc}")


def _b(s: str) -> bytes:
    return os.fsencode(s)


# ---------------------------------------------------------------------------------------------------- <idx>_feature
def feature_path(fold: str, idx: int) -> str:
    buf = C.create_string_buffer(4096)
    _chk(lib().msfm_feature_path(_b(fold), idx, buf, 4096), "msfm_feature_path")
    return os.fsdecode(buf.value)


def feature_stat(path: str) -> FeatureInfo:
    info = FeatureInfo()
    _chk(lib().msfm_feature_stat(_b(path), C.byref(info)), f"msfm_feature_stat({path})")
    return info


def feature_read(path: str):
    """Returns dict(info, maker, model, xy [n,2] float32 centred, desc [rows, cols] float32 or uint8)."""
    info = feature_stat(path)
    maker = C.create_string_buffer(info.maker_len + 1)
    model = C.create_string_buffer(info.model_len + 1)
    xy = np.empty((info.num_pts, 2), np.float32)
    dtype = {5: np.float32, 0: np.uint8}.get(info.desc_type)
    if dtype is None:
        raise StoreError(f"unsupported descriptor type code {info.desc_type}")
    desc = np.empty((info.desc_rows, info.desc_cols), dtype)
    _chk(lib().msfm_feature_read(_b(path), C.byref(info), maker, model, xy.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p),
                                 C.c_int64(desc.strides[0] if desc.size else info.desc_cols * info.desc_elem_size)), "msfm_feature_read")
    return dict(info=info, maker=maker.value.decode("latin-1"), model=model.value.decode("latin-1"), xy=xy, desc=desc)


def feature_write(path: str, *, rows: int, cols: int, xy_pixel: np.ndarray, desc: np.ndarray, zoom_ratio=1.0, f_mm=0.0, f_pixel=0.0,
                  gps_latitude=0.0, gps_longitude=0.0, maker="", model="") -> None:
    xy_pixel = np.ascontiguousarray(xy_pixel, np.float32).reshape(-1, 2)
    if desc.dtype not in (np.float32, np.uint8):
        raise TypeError("descriptors must be float32 or uint8")
    desc = np.ascontiguousarray(desc)
    info = FeatureInfo()
    info.rows, info.cols = rows, cols
    info.zoom_ratio, info.f_mm, info.f_pixel, info.gps_latitude, info.gps_longitude = zoom_ratio, f_mm, f_pixel, gps_latitude, gps_longitude
    info.num_pts = xy_pixel.shape[0]
    info.desc_rows, info.desc_cols = desc.shape
    info.desc_type = 5 if desc.dtype == np.float32 else 0
    _chk(lib().msfm_feature_write(_b(path), C.byref(info), maker.encode("latin-1"), model.encode("latin-1"),
                                  xy_pixel.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p),
                                  C.c_int64(desc.strides[0] if desc.size else desc.shape[1] * desc.itemsize)), "msfm_feature_write")


# ---------------------------------------------------------------------------------------------------- <idx1>_match
def match_append(fold: str, idx1: int, idx2: int, pairs: np.ndarray) -> None:
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    _chk(lib().msfm_match_append(_b(fold), idx1, idx2, pairs.ctypes.data_as(C.c_void_p), pairs.shape[0]), "msfm_match_append")


def match_read(fold: str, idx1: int):
    """Graph::QueryMatch(idx, image_ids, match_pts): returns (image_ids [r] int32, [pairs [n,2] int32 per record])."""
    n_rec, n_pairs = C.c_int32(), C.c_int64()
    rc = lib().msfm_match_read(_b(fold), idx1, None, None, 0, None, C.c_int64(0), C.byref(n_rec), C.byref(n_pairs))
    if rc not in (0, ERR_CAPACITY):
        raise StoreError(f"msfm_match_read failed with status {rc}")
    ids = np.empty((n_rec.value,), np.int32)
    offs = np.zeros((n_rec.value + 1,), np.int64)
    pairs = np.empty((max(n_pairs.value, 1), 2), np.int32)
    _chk(lib().msfm_match_read(_b(fold), idx1, ids.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), n_rec.value,
                               pairs.ctypes.data_as(C.c_void_p), C.c_int64(n_pairs.value), C.byref(n_rec), C.byref(n_pairs)), "msfm_match_read")
    return ids, [pairs[offs[r]:offs[r + 1]].copy() for r in range(n_rec.value)]


def match_index_missing(fold: str, num_imgs: int) -> np.ndarray:
    out = np.empty((max(num_imgs, 1),), np.int32)
    n = C.c_int32()
    _chk(lib().msfm_match_index_missing(_b(fold), num_imgs, out.ctypes.data_as(C.c_void_p), C.byref(n)), "msfm_match_index_missing")
    return out[:n.value].copy()


def match_index_append(fold: str, idx1: int) -> None:
    _chk(lib().msfm_match_index_append(_b(fold), idx1), "msfm_match_index_append")


def graph_write(fold: str, graph: np.ndarray) -> None:
    g = np.ascontiguousarray(graph, np.int32)
    _chk(lib().msfm_graph_write(_b(fold), g.shape[0], g.ctypes.data_as(C.c_void_p)), "msfm_graph_write")


def graph_read(fold: str, num_imgs: int) -> np.ndarray:
    g = np.empty((num_imgs, num_imgs), np.int32)
    _chk(lib().msfm_graph_read(_b(fold), num_imgs, g.ctypes.data_as(C.c_void_p)), "msfm_graph_read")
    return g


def graph_recover(fold: str, num_imgs: int, existing) -> np.ndarray:
    e = np.ascontiguousarray(existing, np.int32)
    g = np.empty((num_imgs, num_imgs), np.int32)
    _chk(lib().msfm_graph_recover(_b(fold), num_imgs, e.ctypes.data_as(C.c_void_p), e.shape[0], g.ctypes.data_as(C.c_void_p)), "msfm_graph_recover")
    return g


# ---------------------------------------------------------------------------------------------------- pair lists
def _adjacency(offsets: np.ndarray, lst: np.ndarray):
    return [lst[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]


def pairs_all(num_imgs: int):
    offs = np.zeros((num_imgs + 1,), np.int64)
    lst = np.empty((max(num_imgs * (num_imgs - 1), 1),), np.int32)
    _chk(lib().msfm_pairs_all(num_imgs, offs.ctypes.data_as(C.c_void_p), lst.ctypes.data_as(C.c_void_p)), "msfm_pairs_all")
    return offs, lst[:offs[-1]]


def pairs_priori_xy(xy: np.ndarray, knn: int = 50):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    n = xy.shape[0]
    offs = np.zeros((n + 1,), np.int64)
    lst = np.empty((max(n * max(min(knn, n // 10), 0), 1),), np.int32)
    _chk(lib().msfm_pairs_priori_xy(n, xy.ctypes.data_as(C.c_void_p), knn, offs.ctypes.data_as(C.c_void_p), lst.ctypes.data_as(C.c_void_p)),
         "msfm_pairs_priori_xy")
    return offs, lst[:offs[-1]]


def _words_csr(words_per_image):
    offs = np.zeros((len(words_per_image) + 1,), np.int64)
    for i, w in enumerate(words_per_image):
        offs[i + 1] = offs[i] + len(w)
    flat = np.concatenate([np.asarray(w, np.int32).reshape(-1) for w in words_per_image]) if len(words_per_image) and offs[-1] else np.zeros((0,), np.int32)
    return offs, np.ascontiguousarray(flat, np.int32)


def similarity_invfile(words_per_image, num_words: int) -> np.ndarray:
    """SimilarityGraph::SimilarityGraphInvFile (similarity_graph.cc:47-117): [n, n] float32 counts of shared words."""
    offs, flat = _words_csr(words_per_image)
    n = len(words_per_image)
    sim = np.zeros((n, n), np.float32)
    _chk(lib().msfm_similarity_invfile(n, offs.ctypes.data_as(C.c_void_p), flat.ctypes.data_as(C.c_void_p), int(num_words),
                                       sim.ctypes.data_as(C.c_void_p)), "msfm_similarity_invfile")
    return sim


def pairs_similarity_topk(similarity: np.ndarray, th_num_match: int = 0):
    """Hypotheses of InitialMatchingGraph::match_graph_feature (initial_matching_graph.cc:166-168, 212-231)."""
    sim = np.ascontiguousarray(similarity, np.float32)
    n = sim.shape[0]
    k = th_num_match if th_num_match > 0 else min(min(max(200, n // 10), n - 1), 500)
    offs = np.zeros((n + 1,), np.int64)
    lst = np.empty((max(n * max(k, 0), 1),), np.int32)
    _chk(lib().msfm_pairs_similarity_topk(n, sim.ctypes.data_as(C.c_void_p), int(th_num_match), offs.ctypes.data_as(C.c_void_p),
                                          lst.ctypes.data_as(C.c_void_p)), "msfm_pairs_similarity_topk")
    return offs, lst[:offs[-1]]


def word_matches(words1, words2) -> np.ndarray:
    """Word-collision matches (pt1, pt2) of two images (initial_matching_graph.cc:239-251)."""
    w1 = np.ascontiguousarray(words1, np.int32).reshape(-1)
    w2 = np.ascontiguousarray(words2, np.int32).reshape(-1)
    cap = max(min(len(w1), len(w2)), 1)
    out = np.empty((cap, 2), np.int32)
    n = lib().msfm_word_matches(w1.ctypes.data_as(C.c_void_p), len(w1), w2.ctypes.data_as(C.c_void_p), len(w2), out.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        raise StoreError(f"msfm_word_matches failed with status {n}")
    return out[:n].copy()


def init_graph_write(fold: str, offsets: np.ndarray, lst: np.ndarray, id_last: int) -> None:
    offsets = np.ascontiguousarray(offsets, np.int64)
    lst = np.ascontiguousarray(lst, np.int32)
    _chk(lib().msfm_init_graph_write(_b(fold), len(offsets) - 1, id_last, offsets.ctypes.data_as(C.c_void_p), lst.ctypes.data_as(C.c_void_p)),
         "msfm_init_graph_write")


def init_graph_read(fold: str):
    n, id_last, n_list = C.c_int32(), C.c_int32(), C.c_int64()
    rc = lib().msfm_init_graph_read(_b(fold), C.byref(n), C.byref(id_last), None, 0, None, C.c_int64(0), C.byref(n_list))
    if rc not in (0, ERR_CAPACITY):
        raise StoreError(f"msfm_init_graph_read failed with status {rc}")
    offs = np.zeros((n.value + 1,), np.int64)
    lst = np.empty((max(n_list.value, 1),), np.int32)
    _chk(lib().msfm_init_graph_read(_b(fold), C.byref(n), C.byref(id_last), offs.ctypes.data_as(C.c_void_p), n.value + 1,
                                    lst.ctypes.data_as(C.c_void_p), C.c_int64(n_list.value), C.byref(n_list)), "msfm_init_graph_read")
    return offs, lst[:n_list.value], id_last.value


# ---------------------------------------------------------------------------------------------------- driver
def build_match_graph(fold: str, offsets: np.ndarray, lst: np.ndarray, *, device=0, th_good=0.6, th_all=0.85, mutual=False,
                      min_keypoints=0, min_good=0, descriptor_scale=1.0, rescore_band=0.0, verify=None, geo_verify=False,
                      geo_seed=0, max_batch_rows=0, n_devices=1) -> None:
    """FineMatchingGraph::BuildMatchGraph on the GPU matcher (see include/msfm_graph.h).  `verify(idx1, idx2, xy1, xy2,
    matches, good)` -> (accept, keep_indices) is the geo-verification seam."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    lst = np.ascontiguousarray(lst, np.int32)
    opt = GraphOptions(device, th_good, th_all, int(bool(mutual)), min_keypoints, min_good, descriptor_scale, rescore_band,
                       int(bool(geo_verify)), geo_seed, max_batch_rows, int(n_devices), 0)
    cb = None
    if verify is not None:
        def _trampoline(_user, idx1, idx2, xy1, n1, xy2, n2, m, g, n, keep, n_keep):
            a1 = np.ctypeslib.as_array(xy1, shape=(n1, 2)) if n1 else np.empty((0, 2), np.float32)
            a2 = np.ctypeslib.as_array(xy2, shape=(n2, 2)) if n2 else np.empty((0, 2), np.float32)
            mm = np.ctypeslib.as_array(m, shape=(n, 2)) if n else np.empty((0, 2), np.int32)
            gg = np.ctypeslib.as_array(g, shape=(n,)) if n else np.empty((0,), np.uint8)
            accept, idx = verify(idx1, idx2, a1, a2, mm, gg)
            idx = np.asarray(idx, np.int32)
            for k, v in enumerate(idx):
                keep[k] = int(v)
            n_keep[0] = len(idx)
            return int(bool(accept))
        cb = VERIFY_FN(_trampoline)
    err = C.create_string_buffer(512)
    rc = graph_lib().msfm_build_match_graph(_b(fold), len(offsets) - 1, offsets.ctypes.data_as(C.c_void_p), lst.ctypes.data_as(C.c_void_p),
                                            C.byref(opt), cb if cb is not None else None, None, err, 512)
    if rc != 0:
        raise StoreError(f"msfm_build_match_graph failed ({rc}): {err.value.decode()}")
