#!/bin/bash
# Throughput of the hot path at the other BASELINE config shapes (device-resident, mutual on), one JSON line each.
for cfg in "20000 30" "16384 30" "32768 20"; do set -- $cfg
  python bench.py --rows $1 --images $2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-int8-peak 2>/dev/null
done
