#!/bin/bash
# Quick GPU check used during kernel work: parity tests, one-way / mutual throughput on 40 images, debug experiments.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --images 40 --no-cpu-baseline --no-e2e --mutual $1 --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('pairs/s', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms_per_step'],3), 'step_ms', round(d['ms_per_step'],3))
    else: print(l)
"; }
echo "== one-way"; run 0
echo "== mutual"; run 1
for f in ${DEBUG_FLAGS:-8 1 2}; do echo "== MSFM_DEBUG_FLAGS=$f (one-way)"; MSFM_DEBUG_FLAGS=$f run 0; done
