"""Python face of include/msfm_multi.h — the single-process multi-GPU pair scheduler (north_star subsystem 4).

Nothing is computed here: staging (H2D on the owner device + NCCL broadcast over NVLink into the other devices' tables),
sharding (msfm_sched_shard), per-device matching threads and the stitching of the match lists into one result in the
caller's pair order all happen behind the C ABI in libmsfm_match.so (csrc/msfm_multi.cc).  This is what a MetricSfM
build — one C++ process, FineMatchingGraph::BuildMatchGraph, SfM/src/graph/fine_matching_graph.cc:58-133 — would call
to use every GPU of the box; tests and bench.py drive it through this mirror.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MultiConfig, MultiTiming, Result, Timing
from .matcher import Matcher, MatchResult, MsfmError, _host_ptr


class MultiMatcher:
    def __init__(self, devices, max_images: int, arena_rows: int):
        self._L = _lib.load()
        devs = np.ascontiguousarray(devices, np.int32)
        cfg = MultiConfig()
        cfg.n_devices = len(devs)
        cfg.devices = devs.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.max_images = max_images
        cfg.arena_rows = arena_rows
        h = C.c_void_p()
        st = self._L.msfm_multi_create(C.byref(cfg), C.byref(h))
        if st != _lib.MSFM_OK:
            raise MsfmError(st, self._L.msfm_status_string(st).decode())
        self._h = h
        self.n_devices = len(devs)
        self._inflight = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.msfm_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != _lib.MSFM_OK:
            raise MsfmError(st, self._L.msfm_multi_last_error(self._h).decode())

    def _marshal(self, image_ids, descs, dtype_name):
        n = len(image_ids)
        ids = np.ascontiguousarray(image_ids, np.int32)
        ptrs = (C.c_void_p * max(n, 1))()
        rows = np.zeros((max(n, 1),), np.int32)
        for k, d in enumerate(descs):
            is_t = hasattr(d, "numpy") and not isinstance(d, np.ndarray)
            if not is_t:
                d = np.asarray(d)
            dense = d.is_contiguous() if is_t else d.flags["C_CONTIGUOUS"]
            if str(d.dtype).replace("torch.", "") != dtype_name or len(d.shape) != 2 or (d.shape[0] and d.shape[1] != 128) or not dense:
                raise ValueError(f"descriptors must be dense [rows, 128] {dtype_name}")
            rows[k] = d.shape[0]
            ptr, keep = _host_ptr(d)
            ptrs[k] = ptr
            self._inflight.append(keep)
        return n, ids, ptrs, rows

    def upload_u8(self, image_ids, descs) -> None:
        """Stage a group of uint8 images: packed on their owner device, broadcast to the others.  No host wait."""
        n, ids, ptrs, rows = self._marshal(image_ids, descs, "uint8")
        self._check(self._L.msfm_multi_upload_u8(self._h, n, ids.ctypes.data, C.cast(ptrs, C.c_void_p), rows.ctypes.data))

    def upload_f32(self, image_ids, descs, scale: float = 1.0) -> None:
        n, ids, ptrs, rows = self._marshal(image_ids, descs, "float32")
        self._check(self._L.msfm_multi_upload_f32(self._h, n, ids.ctypes.data, C.cast(ptrs, C.c_void_p), rows.ctypes.data, scale))

    def sync(self) -> None:
        self._check(self._L.msfm_multi_sync(self._h))
        self._inflight = []

    def release_all(self) -> None:
        self._check(self._L.msfm_multi_release_all(self._h))
        self._inflight = []

    def match_pairs(self, pairs, ratio: float = 0.6, *, ratio_good: float = 0.0, max_dist_sq: float = 0.0, mutual: bool = False,
                    min_keypoints: int = 20, orientation: int = 0, capacity: int, out: MatchResult | None = None,
                    flags: int = 0) -> MatchResult:
        """msfm_multi_match_pairs; `capacity` = entries of the result buffers (at most one match per query row)."""
        pa = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pa.shape[0]
        if out is None:
            out = MatchResult(offsets=np.zeros((n + 1,), np.int64), ok=np.zeros((max(n, 1),), np.int32),
                              matches=np.empty((max(capacity, 1), 2), np.int32),
                              good=np.zeros((max(capacity, 1),), np.uint8) if ratio_good > 0 else None)
        res = Result()
        res.offsets = out.offsets.ctypes.data_as(_lib._i64p)
        res.ok = out.ok.ctypes.data_as(_lib._i32p)
        res.matches = _host_ptr(out.matches)[0]
        res.good = out.good.ctypes.data_as(_lib._u8p) if out.good is not None else None
        res.match_capacity = out.matches.shape[0]
        prm = Matcher._params(ratio, ratio_good, max_dist_sq, mutual, min_keypoints, orientation, 0.0, flags)
        self._check(self._L.msfm_multi_match_pairs(self._h, pa.ctypes.data, n, C.byref(prm), C.byref(res)))
        total = int(out.offsets[n])
        return MatchResult(out.offsets, out.ok[:n], out.matches[:total], out.good[:total] if out.good is not None else None)

    def timing(self):
        t = MultiTiming()
        per = (Timing * self.n_devices)()
        self._check(self._L.msfm_multi_last_timing(self._h, C.byref(t), C.cast(per, C.c_void_p)))
        return ({k: getattr(t, k) for k, _ in MultiTiming._fields_},
                [{k: getattr(p, k) for k, _ in Timing._fields_} for p in per])

    def download_packed(self, device_slot: int, image_id: int):
        """The packed rows of an image as device `device_slot` holds them (owner-packed or NCCL-received)."""
        self._check(self._L.msfm_multi_sync(self._h))   # uploads and broadcasts have landed
        ctx = self._L.msfm_multi_context(self._h, device_slot)
        rows, off = C.c_int32(), C.c_int64()
        st = self._L.msfm_image_info(ctx, image_id, C.byref(rows), C.byref(off))
        if st != _lib.MSFM_OK:
            raise MsfmError(st, self._L.msfm_last_error(ctx).decode())
        desc = np.empty((rows.value, 128), np.uint8)
        norms = np.empty((rows.value,), np.uint32)
        st = self._L.msfm_download_packed(ctx, image_id, desc.ctypes.data, norms.ctypes.data)
        if st != _lib.MSFM_OK:
            raise MsfmError(st, self._L.msfm_last_error(ctx).decode())
        return desc, norms
