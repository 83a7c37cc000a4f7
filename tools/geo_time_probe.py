import sys, time, numpy as np
sys.path.insert(0, '.')
from metricsfm_b200 import synth
from metricsfm_b200.matcher import Matcher
n=30; rows=8192
col=synth.Collection(rows, seed=0)
imgs=[col.image_u8(i) for i in range(n)]
pairs=synth.exhaustive_pairs(n)
rng=np.random.default_rng(0)
xy={i: rng.uniform(-2000,2000,(rows,2)).astype(np.float32) for i in range(n)}
with Matcher(device=0, max_images=n, arena_rows=n*rows) as m:
    m.upload_batch(list(range(n)), imgs)
    res=m.match_pairs(pairs, 0.85, ratio_good=0.6, mutual=True)
    print('pairs', len(pairs), 'matches', len(res.matches), 'good', int(res.good.sum()))
    for iters in (1024, 256):
        t0=time.perf_counter()
        ok,inl,keep,F=m.geo_verify(pairs,res,xy,iters=iters)
        t1=time.perf_counter()
        print('iters',iters,'wall_s',round(t1-t0,4),'kernel_ms',m.timing()['finalize_ms'],'ok',int(ok.sum()))
