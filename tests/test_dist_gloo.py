"""World-size-2 gloo test (CPU) of the multi-GPU host path: owner blocks, table replication by broadcast, cost-based pair
sharding, per-rank matching, gather + stitch.  The per-rank matcher is the CPU oracle here (tests may use it); on GPUs
bench.py runs the same functions with the CUDA library and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_IMAGES, ROWS = 6, 300
ROWS_PADDED = 512


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    from metricsfm_b200 import distributed as D, synth
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    col = synth.Collection(ROWS, seed=77)
    rows = np.array([ROWS - 7 * i for i in range(N_IMAGES)])
    owner, ranges = D.block_ranges(N_IMAGES, ROWS_PADDED, world)
    desc = torch.zeros((N_IMAGES * ROWS_PADDED, 128), dtype=torch.uint8)
    side = torch.zeros((N_IMAGES * ROWS_PADDED,), dtype=torch.int32)
    for gid in range(N_IMAGES):
        if owner[gid] == rank:                      # "pack + upload" only what this rank owns
            d = col.image_u8(gid, rows[gid])
            desc[gid * ROWS_PADDED: gid * ROWS_PADDED + rows[gid]] = torch.from_numpy(d)
            side[gid * ROWS_PADDED: gid * ROWS_PADDED + rows[gid]] = torch.from_numpy((d.astype(np.int64) ** 2).sum(1).astype(np.int32))
    got = D.replicate_arena(desc, side, ranges, dist)
    assert got == sum((hi - lo) * 132 for r, (lo, hi) in enumerate(ranges) if r != rank)
    images = [desc[g * ROWS_PADDED: g * ROWS_PADDED + rows[g]].numpy() for g in range(N_IMAGES)]
    for gid in range(N_IMAGES):                     # every rank now holds the whole table
        np.testing.assert_array_equal(images[gid], col.image_u8(gid, rows[gid]))
    pairs = synth.exhaustive_pairs(N_IMAGES)

    def match_fn(local_pairs):
        offsets, ok, chunks, goods = [0], [], [], []
        for r, q in local_pairs:
            res = oracle.match_pair_u8(images[r], images[q], 0.85, mutual=True, ratio_good=0.6)
            ok.append(1 if res["ok"] else 0)
            chunks.append(res["pairs"])
            goods.append(res["good"])
            offsets.append(offsets[-1] + len(res["pairs"]))
        return dict(offsets=np.array(offsets, np.int64), ok=np.array(ok, np.int32),
                    matches=np.concatenate(chunks) if chunks else np.zeros((0, 2), np.int32),
                    good=np.concatenate(goods) if goods else np.zeros((0,), np.uint8))

    shards, local = D.match_sharded(match_fn, pairs, rows, rank, world)
    assert sorted(np.concatenate(shards).tolist()) == list(range(len(pairs)))
    stitched = D.gather_results(len(pairs), shards, local, dist)
    if rank == 0:
        offsets, ok, matches, good = stitched
        np.savez(os.path.join(out_dir, "stitched.npz"), offsets=offsets, ok=ok, matches=matches, good=good)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replication_sharding_and_gather(tmp_path, oracle_mod):
    from metricsfm_b200 import synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "stitched.npz")
    col = synth.Collection(ROWS, seed=77)
    rows = [ROWS - 7 * i for i in range(N_IMAGES)]
    imgs = [col.image_u8(g, rows[g]) for g in range(N_IMAGES)]
    pairs = synth.exhaustive_pairs(N_IMAGES)
    off = 0
    for p, (r, q) in enumerate(pairs):
        exp = oracle_mod.match_pair_u8(imgs[r], imgs[q], 0.85, mutual=True, ratio_good=0.6)
        n = len(exp["pairs"])
        assert got["offsets"][p] == off and got["ok"][p] == 1
        np.testing.assert_array_equal(got["matches"][off:off + n], exp["pairs"])
        np.testing.assert_array_equal(got["good"][off:off + n], exp["good"])
        off += n
    assert got["offsets"][-1] == off and off > 0
