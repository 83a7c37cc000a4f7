"""Time the pieces of the multi-GPU table staging (run under torchrun): H2D of the own block, the two in-place NCCL
all-gathers (descriptors, keys), each bracketed by device synchronisation."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metricsfm_b200 import distributed as D
from metricsfm_b200.matcher import Matcher

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_local, rows = 100, 8192
n_global = n_local * world
dev = torch.device("cuda", lr)
desc = torch.empty((n_global * rows, 128), dtype=torch.uint8, device=dev)
keys = torch.empty((n_global * rows,), dtype=torch.int32, device=dev)
host = torch.randint(0, 255, (n_local, rows, 128), dtype=torch.uint8).pin_memory()
m = Matcher(device=lr, max_images=n_global, arena_rows=n_global * rows, external_desc_arena=desc.data_ptr(), external_norm_arena=keys.data_ptr())
owner, ranges = D.block_ranges(n_global, rows, world)
mine = [g for g in range(n_global) if owner[g] == rank]
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.release_all()
    before = [g for g in range(n_global) if g < mine[0]]; after = [g for g in range(n_global) if g > mine[-1]]
    if before: m.reserve_batch(before, [rows] * len(before))
    m.upload_batch(mine, [host[g - mine[0]] for g in mine], wait=False)
    if after: m.reserve_batch(after, [rows] * len(after))
    t1 = time.perf_counter()
    m.sync()
    t2 = time.perf_counter()
    lo, hi = ranges[rank]
    dist.all_gather_into_tensor(desc, desc[lo:hi]); torch.cuda.synchronize()
    t3 = time.perf_counter()
    dist.all_gather_into_tensor(keys, keys[lo:hi]); torch.cuda.synchronize()
    t4 = time.perf_counter()
    dist.barrier(); torch.cuda.synchronize()
    t5 = time.perf_counter()
    if rank == 0:
        print(f"iter {it}: enqueue {1e3*(t1-t0):.2f} | H2D wait {1e3*(t2-t1):.2f} | gather desc {1e3*(t3-t2):.2f} ({desc.numel()/1e6:.0f} MB) | "
              f"gather keys {1e3*(t4-t3):.2f} | barrier {1e3*(t5-t4):.2f} ms", flush=True)
dist.destroy_process_group()
