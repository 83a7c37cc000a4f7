#!/bin/bash
# A/B of kernel variants on one box: for every library in exp_libs/ (built by tools/variants.sh build ...) run the headline
# bench (mutual and one-way).  usage: bash tools/ab_bench.sh [extra bench args]
cd "$(dirname "$0")/.."
cp metricsfm_b200/csrc/libmsfm_match.so /tmp/orig.so
for f in exp_libs/*.so; do
    n=$(basename $f .so); cp $f metricsfm_b200/csrc/libmsfm_match.so
    for m in ${AB_MODES:-1 0}; do
        timeout 300 python bench.py --no-cpu-baseline --no-int8-peak --mutual $m --steps 5 --warmup 3 "$@" > gpurun_out/ab_${n}_m$m.json 2> gpurun_out/ab_${n}_m$m.err
        python - gpurun_out/ab_${n}_m$m.json $n $m <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); e=d.get("e2e") or {}
    print(f"{sys.argv[2]:28s} mutual={sys.argv[3]} value {d['value']:.0f} kernel_ms {d['roofline']['kernel_ms_per_step']:.2f} step_ms {d['ms_per_step']:.2f} e2e {e.get('value',0):.0f} (u8 {e.get('uint8_rows',{}).get('value',0):.0f}) parity {d.get('parity_ok')} twin {d['config'].get('mutual_pairs_needing_tensor_twin_pass_rank0')} clk {d['clocks'].get('sm_mhz')}")
except Exception as ex: print(sys.argv[2], "FAILED", ex)
PY
    done
done
cp /tmp/orig.so metricsfm_b200/csrc/libmsfm_match.so
