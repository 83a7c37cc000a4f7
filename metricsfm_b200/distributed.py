"""Multi-GPU plumbing of the pair scheduler (north_star subsystem 4): one process per GPU, torch.distributed.

  1. every rank packs + uploads its own block of images into its arena, reserves the others in the same order
     (identical offsets on every rank),
  2. replicate_arena(): the packed table (u8 rows + per-row side words) is replicated with one broadcast per owner block
     — NCCL over NVLink/NVSwitch on GPUs, gloo on CPU in the tests,
  3. every rank matches its own shard of the pair list (scheduler.shard_pairs) — no data-path collective,
  4. gather_results(): per-rank match lists go back to rank 0 and are stitched into global pair order.

The functions take tensors / an abstract `match_fn`, so the same code path is exercised by tests/test_dist_gloo.py on
CPU (gloo, world_size 2) and by bench.py on GPUs (nccl).
"""
from __future__ import annotations

import os

import numpy as np

from . import scheduler


def pin_to_gpu_numa_node(device_index: int) -> dict:
    """Restrict this process to the CPU cores NVML reports as local to GPU `device_index`, so that the page-locked
    staging buffers allocated afterwards land on that NUMA node and host->device copies do not cross sockets (with one
    process per GPU, eight ranks otherwise share whatever node the scheduler put them on).  Returns what was done;
    never raises (a missing NVML or a refused affinity call leaves the process unchanged)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cores & allowed
        if not target:
            return {"pinned": False, "why": "no overlap between NVML affinity and the allowed cores"}
        if target != allowed:
            os.sched_setaffinity(0, target)
        return {"pinned": True, "cores": len(target), "of": len(allowed)}
    except Exception as exc:  # noqa: BLE001
        return {"pinned": False, "why": str(exc)[:120]}


def block_ranges(n_images: int, rows_padded_per_image, world_size: int):
    """Arena row range [lo, hi) owned by each rank when images are allocated in id order, owner = contiguous blocks."""
    owner = scheduler.image_owner(n_images, world_size)
    rp = np.broadcast_to(np.asarray(rows_padded_per_image, dtype=np.int64), (n_images,))
    starts = np.concatenate([[0], np.cumsum(rp)])
    ranges = []
    for r in range(world_size):
        ids = np.nonzero(owner == r)[0]
        ranges.append((int(starts[ids[0]]), int(starts[ids[-1] + 1])) if len(ids) else (0, 0))
    return owner, ranges


def replicate_arena(desc_arena, side_arena, ranges, dist) -> int:
    """Replicate every owner's block of the packed table on all ranks.  Returns the bytes this rank received.

    The collective is chosen from facts every rank agrees on (backend name, block layout) BEFORE anything is issued, so all
    ranks always issue the same sequence; errors of the collectives themselves propagate."""
    rank = dist.get_rank()
    world = dist.get_world_size()
    row_bytes = desc_arena.shape[1] * desc_arena.element_size() + side_arena.element_size()
    # Equal, contiguous owner blocks in rank order (the usual case: equally sized images) on NCCL: ONE in-place all-gather
    # per arena over NVLink/NVSwitch (the input is this rank's own slice of the output).
    block = ranges[0][1] - ranges[0][0]
    equal_blocks = block > 0 and all(r == (i * block + ranges[0][0], (i + 1) * block + ranges[0][0]) for i, r in enumerate(ranges))
    if world > 1 and equal_blocks and str(dist.get_backend()).lower() == "nccl":
        base = ranges[0][0]
        lo, hi = ranges[rank]
        dist.all_gather_into_tensor(desc_arena[base:base + world * block], desc_arena[lo:hi])
        dist.all_gather_into_tensor(side_arena[base:base + world * block], side_arena[lo:hi])
        return (world - 1) * block * row_bytes
    received = 0
    for src, (lo, hi) in enumerate(ranges):     # ragged blocks, or a backend without in-place all-gather (gloo in the CPU tests)
        if hi <= lo:
            continue
        dist.broadcast(desc_arena[lo:hi], src=src)
        dist.broadcast(side_arena[lo:hi], src=src)
        if src != rank:
            received += (hi - lo) * row_bytes
    return received


def match_sharded(match_fn, pairs, rows_per_image, rank: int, world_size: int):
    """Shard the pair list by cost and run `match_fn(local_pairs) -> dict(offsets, ok, matches, good)` on this rank's
    shard.  Returns (shard index arrays of every rank, this rank's result)."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    shards = scheduler.shard_pairs(pairs, rows_per_image, world_size)
    return shards, match_fn(pairs[shards[rank]])


def gather_results(n_pairs: int, shards, local_result, dist, dst: int = 0):
    """Collect the per-rank match lists on `dst` and stitch them back into global pair order.
    Returns (offsets, ok, matches, good) on dst, None elsewhere."""
    world = dist.get_world_size()
    payload = {k: (np.asarray(v) if v is not None else None) for k, v in local_result.items()}
    gathered = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    return scheduler.stitch_results(n_pairs, shards, gathered)
