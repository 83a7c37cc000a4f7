"""Where the end-to-end step spends its host time (1 GPU): upload_batch, match_pairs (host wall vs device total), D2H."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metricsfm_b200 import synth
from metricsfm_b200.matcher import Matcher, MatchResult

n, rows = 100, 8192
col = synth.Collection(rows, seed=0)
host = torch.empty((n, rows, 128), dtype=torch.uint8).pin_memory()
for i in range(n):
    host[i].numpy()[:] = col.image_u8(i)
pairs = synth.exhaustive_pairs(n)
m = Matcher(device=0, max_images=n, arena_rows=n * rows)
cap = len(pairs) * 2048
out = MatchResult(offsets=np.zeros((len(pairs) + 1,), np.int64), ok=np.zeros((len(pairs),), np.int32),
                  matches=torch.empty((cap, 2), dtype=torch.int32).pin_memory().numpy(),
                  good=torch.empty((cap,), dtype=torch.uint8).pin_memory().numpy())
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.release_all()
    m.upload_batch(list(range(n)), [host[i] for i in range(n)])
    t1 = time.perf_counter()
    res = m.match_pairs(pairs, 0.85, ratio_good=0.6, mutual=True, min_keypoints=20, out=out)
    t2 = time.perf_counter()
    t = m.timing()
    print(f"iter {it}: upload {1e3*(t1-t0):.2f} ms | match_pairs wall {1e3*(t2-t1):.2f} ms, device total {t['total_ms']:.2f}, "
          f"kernel {t['match_kernel_ms']:.2f}, finalize {t['finalize_ms']:.2f}, d2h {t['d2h_ms']:.2f} ms ({t['d2h_bytes']/1e6:.1f} MB)")
