"""Pair scheduler (north_star subsystem 4): shards a candidate pair list across the GPUs of one box.

Pairs are independent units (the reference already treats them so: the OpenMP partner loop at
SfM/src/graph/fine_matching_graph.cc:87), so there is no data-path collective during matching: the packed descriptor
table is replicated once (NCCL over NVLink, see replicate_table), every rank matches its own shard, and the match
lists go back to the host per rank.

Pure host logic (numpy only) so it is testable on CPU with the gloo backend.
"""
from __future__ import annotations

import numpy as np


def pair_costs(pairs: np.ndarray, rows_per_image: np.ndarray) -> np.ndarray:
    """Algorithmic cost of each pair = M * N (x 2 * 128 int8 ops), SURVEY.md §8d."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    r = np.asarray(rows_per_image, dtype=np.int64)
    return r[p[:, 0]] * r[p[:, 1]]


def shard_pairs(pairs: np.ndarray, rows_per_image: np.ndarray, world_size: int):
    """Greedy longest-processing-time partition of *groups* of pairs into world_size shards.

    Pairs that share the reference image stay together (they are adjacent in the reference's own iteration order,
    fine_matching_graph.cc:58-64, and re-use the same L2-resident reference tiles); groups are assigned, largest first,
    to the currently lightest rank; oversized groups are split so that no group exceeds 1/(4*world) of the total.
    Returns a list of index arrays (positions into `pairs`), one per rank; within a rank the original order is kept.
    """
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    n = p.shape[0]
    if world_size <= 1:
        return [np.arange(n, dtype=np.int64)]
    cost = pair_costs(p, rows_per_image).astype(np.float64)
    total = float(cost.sum())
    cap = max(total / (4.0 * world_size), float(cost.max()) if n else 0.0)
    # groups = runs of equal reference id, cut when a run grows past `cap`
    groups = []
    start, acc = 0, 0.0
    for i in range(n):
        new_run = i > start and p[i, 0] != p[i - 1, 0]
        if i > start and (new_run or acc + cost[i] > cap):
            groups.append((start, i, acc))
            start, acc = i, 0.0
        acc += cost[i]
    if n > start:
        groups.append((start, n, acc))
    order = sorted(range(len(groups)), key=lambda g: -groups[g][2])
    load = np.zeros(world_size)
    owner = [[] for _ in range(world_size)]
    for g in order:
        r = int(np.argmin(load))
        owner[r].append(g)
        load[r] += groups[g][2]
    shards = []
    for r in range(world_size):
        idx = [np.arange(groups[g][0], groups[g][1], dtype=np.int64) for g in sorted(owner[r])]
        shards.append(np.concatenate(idx) if idx else np.zeros((0,), np.int64))
    return shards


def image_owner(n_images: int, world_size: int) -> np.ndarray:
    """Which rank packs + uploads which image before replication: contiguous blocks (one arena range per rank)."""
    per = (n_images + world_size - 1) // world_size
    return np.minimum(np.arange(n_images) // per, world_size - 1).astype(np.int32)


def stitch_results(n_pairs: int, shard_indices, shard_results):
    """Merge per-rank (offsets, ok, matches[, good]) back into global pair order.

    shard_results[r] = dict(offsets=[k+1], ok=[k], matches=[t,2], good=[t] or None) for the pairs shard_indices[r]."""
    counts = np.zeros((n_pairs,), np.int64)
    ok = np.zeros((n_pairs,), np.int32)
    for idx, res in zip(shard_indices, shard_results):
        if len(idx) == 0:
            continue
        counts[idx] = np.diff(res["offsets"])
        ok[idx] = res["ok"]
    offsets = np.zeros((n_pairs + 1,), np.int64)
    np.cumsum(counts, out=offsets[1:])
    matches = np.empty((int(offsets[-1]), 2), np.int32)
    has_good = any(res.get("good") is not None for res in shard_results)
    good = np.zeros((int(offsets[-1]),), np.uint8) if has_good else None
    for idx, res in zip(shard_indices, shard_results):
        so = res["offsets"]
        for k, p in enumerate(idx):
            a, b = so[k], so[k + 1]
            if b > a:
                matches[offsets[p]:offsets[p] + (b - a)] = res["matches"][a:b]
                if good is not None and res.get("good") is not None:
                    good[offsets[p]:offsets[p] + (b - a)] = res["good"][a:b]
    return offsets, ok, matches, good
