"""Pair scheduler (north_star subsystem 4), Python face of include/msfm_sched.h.

The logic lives in C++ (metricsfm_b200/host/msfm_sched.cc; linked into libmsfm_match.so for the single-process multi-GPU
engine of include/msfm_multi.h, and built as the host-only libmsfm_sched.so used here): these functions only marshal
numpy arrays, so the one-process-per-GPU launcher (bench.py under torchrun, metricsfm_b200/distributed.py) and the C++
engine shard and stitch with the same code.

Pairs are independent units (the reference already treats them so: the OpenMP partner loop at
SfM/src/graph/fine_matching_graph.cc:87), so there is no data-path collective during matching: the packed descriptor
table is replicated once (NCCL over NVLink), every rank matches its own shard, the lists go back to the host per rank.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def pair_costs(pairs: np.ndarray, rows_per_image: np.ndarray) -> np.ndarray:
    """Algorithmic cost of each pair = M * N (x 2 * 128 int8 ops), SURVEY.md §8d."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    r = np.asarray(rows_per_image, dtype=np.int64)
    return r[p[:, 0]] * r[p[:, 1]]


def shard_pairs(pairs: np.ndarray, rows_per_image: np.ndarray, world_size: int):
    """msfm_sched_shard: greedy longest-processing-time partition of runs of pairs sharing the reference image.
    Returns a list of index arrays (positions into `pairs`), one per rank; within a rank the original order is kept."""
    p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    rows = np.ascontiguousarray(rows_per_image, dtype=np.int32)
    n = p.shape[0]
    worker = np.zeros((max(n, 1),), np.int32)
    rc = _lib.load_sched().msfm_sched_shard(p.ctypes.data, n, rows.ctypes.data, rows.shape[0], int(world_size), worker.ctypes.data, None)
    if rc != 0:
        raise ValueError("msfm_sched_shard rejected its arguments (image id out of range or world_size < 1)")
    worker = worker[:n]
    return [np.nonzero(worker == r)[0].astype(np.int64) for r in range(world_size)]


def image_owner(n_images: int, world_size: int) -> np.ndarray:
    """msfm_sched_image_owner: which rank packs + uploads which image before replication (contiguous blocks)."""
    owner = np.zeros((max(n_images, 1),), np.int32)
    if _lib.load_sched().msfm_sched_image_owner(int(n_images), int(world_size), owner.ctypes.data) != 0:
        raise ValueError("bad arguments")
    return owner[:n_images]


def stitch_results(n_pairs: int, shard_indices, shard_results):
    """Merge per-rank (offsets, ok, matches[, good]) back into global pair order (msfm_sched_offsets + msfm_sched_scatter).

    shard_results[r] = dict(offsets=[k+1], ok=[k], matches=[t,2], good=[t] or None) for the pairs shard_indices[r]."""
    L = _lib.load_sched()
    counts = np.zeros((n_pairs,), np.int64)
    ok = np.zeros((n_pairs,), np.int32)
    for idx, res in zip(shard_indices, shard_results):
        if len(idx) == 0:
            continue
        counts[idx] = np.diff(np.asarray(res["offsets"], np.int64))
        ok[idx] = res["ok"]
    offsets = np.zeros((n_pairs + 1,), np.int64)
    total = L.msfm_sched_offsets(counts.ctypes.data, n_pairs, offsets.ctypes.data)
    matches = np.empty((max(int(total), 1), 2), np.int32)
    has_good = any(res.get("good") is not None for res in shard_results)
    good = np.zeros((max(int(total), 1),), np.uint8) if has_good else None
    for idx, res in zip(shard_indices, shard_results):
        if len(idx) == 0:
            continue
        pi = np.ascontiguousarray(idx, np.int64)
        lo = np.ascontiguousarray(res["offsets"], np.int64)
        lm = np.ascontiguousarray(res["matches"], np.int32).reshape(-1, 2)
        lg = np.ascontiguousarray(res["good"], np.uint8) if res.get("good") is not None else None
        rc = L.msfm_sched_scatter(pi.ctypes.data, len(pi), lo.ctypes.data, lm.ctypes.data if lm.size else None,
                                  lg.ctypes.data if lg is not None and lg.size else None, offsets.ctypes.data, matches.ctypes.data,
                                  good.ctypes.data if good is not None else None)
        if rc != 0:
            raise ValueError("msfm_sched_scatter rejected its arguments")
    return offsets, ok, matches[:int(total)], (good[:int(total)] if good is not None else None)
