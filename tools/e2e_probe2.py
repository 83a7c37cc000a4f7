"""Where does the end-to-end step spend its time?  Stages 100 x 8192 images (f32 / u8) in 4 groups and matches the
sub-lists, printing host timestamps (ms) of every phase."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metricsfm_b200.matcher import Matcher, MatchResult
from metricsfm_b200.synth_gpu import GpuCollection
from metricsfm_b200 import synth

n, rows, G = 100, 8192, 4
dev = torch.device("cuda", 0)
col = GpuCollection(rows, dev)
u8 = torch.empty((n, rows, 128), dtype=torch.uint8).pin_memory()
for i in range(n):
    u8[i].copy_(col.image_u8(i))
f32 = torch.empty((n, rows, 128), dtype=torch.float32).pin_memory(); f32.copy_(u8)
pairs = synth.exhaustive_pairs(n)
grp = np.arange(n) // (n // G)
pg = np.maximum(grp[pairs[:, 0]], grp[pairs[:, 1]])
order = np.argsort(pg, kind="stable"); pairs = pairs[order]; pg = pg[order]
bounds = [int(np.searchsorted(pg, g)) for g in range(G)] + [len(pairs)]
m = Matcher(device=0, max_images=n, arena_rows=n * rows)
cap = len(pairs) * 2048
out = MatchResult(np.zeros((len(pairs) + G + 1,), np.int64), np.zeros((len(pairs),), np.int32),
                  torch.empty((cap, 2), dtype=torch.int32).pin_memory().numpy(), torch.empty((cap,), dtype=torch.uint8).pin_memory().numpy())
kw = dict(ratio_good=0.6, mutual=True)
def stage(use_f32):
    m.release_all()
    for g in range(G):
        ids = list(range(g * 25, (g + 1) * 25))
        if use_f32: m.upload_f32_batch_async(ids, [f32[i] for i in ids], scale=1.0)
        else: m.upload_batch(ids, [u8[i] for i in ids], wait=False)
for use_f32 in (False, True, False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stage(use_f32); t1 = time.perf_counter(); m.sync(); t2 = time.perf_counter()
    print(f"f32={use_f32}: stage enqueue {1e3*(t1-t0):.2f} ms, staged (sync) {1e3*(t2-t0):.2f} ms")
    r = m.match_pairs_resident(pairs, 0.85, **kw); t3 = time.perf_counter()
    print(f"   resident match of all pairs {1e3*(t3-t2):.2f} ms")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stage(use_f32); ts = [time.perf_counter()]
    done = 0
    for g in range(G):
        a, b = bounds[g], bounds[g + 1]
        sub = MatchResult(out.offsets[a + g:b + g + 1], out.ok[a:b], out.matches[done:], out.good[done:])
        res = m.match_pairs(pairs[a:b], 0.85, out=sub, **kw); done += len(res.matches)
        ts.append(time.perf_counter())
    print("   pipelined: enqueue %.2f ms; sub-lists end at " % (1e3*(ts[0]-t0)) + ", ".join(f"{1e3*(t-t0):.2f}" for t in ts[1:]) + " ms;",
          "pairs per sub-list", [bounds[g+1]-bounds[g] for g in range(G)])
