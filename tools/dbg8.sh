MSFM_DEBUG_FLAGS=8 timeout 200 python bench.py --images 40 --no-cpu-baseline --no-e2e --no-int8-peak --mutual 0 --steps 2 --warmup 1 2>&1 | grep "msfm debug"
timeout 200 python bench.py --images 40 --no-cpu-baseline --no-e2e --no-int8-peak --mutual 0 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): print('one-way pairs/s', round(json.loads(l)['value']))"
