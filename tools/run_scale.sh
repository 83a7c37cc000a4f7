#!/bin/bash
# Multi-GPU measurement campaign (run under `gpurun --gpus N`): usage tools/run_scale.sh N "<workloads>" [single]
# Writes one JSON line per run into gpurun_out/r2_scale_w<workload>_<N>gpu[_single].json (+ .err).
N=$1; WL=${2:-"2 3"}; SINGLE=$3
run() {  # workload, extra args
  local w=$1; shift
  local out=gpurun_out/r2_scale_w${w}_${N}gpu
  if [ "$N" = "1" ]; then python bench.py --gpus 1 --workload $w "$@" > $out.json 2> $out.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w "$@" > $out.json 2> $out.err; fi
  python - "$out.json" <<PY || tail -5 $out.err
import json,sys
d=json.load(open(sys.argv[1])); e=d.get("e2e") or {}
print("w$w N=$N value", round(d["value"]), "pairs/s ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(e.get("value",0)), "parity", d.get("parity_checked"), d.get("parity_ok"), "clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "pairs", d["config"]["pairs_per_step_all_gpus"])
PY
}
for w in $WL; do
  case $w in
    2) run 2 --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak ;;
    3) run 3 --steps 3 --warmup 2 --e2e-steps 3 --no-cpu-baseline --no-int8-peak ;;
    4) run 4 --steps 3 --warmup 2 --e2e-steps 3 --no-cpu-baseline --no-int8-peak ;;
    5) run 5 --steps 2 --warmup 1 --e2e-steps 2 --parity-pairs 3 --no-cpu-baseline --no-int8-peak ;;
  esac
done
if [ -n "$SINGLE" ]; then
  out=gpurun_out/r2_scale_w2_${N}gpu_single
  python bench.py --gpus $N --engine single --workload 2 --steps 5 --warmup 3 > $out.json 2> $out.err
  python - "$out.json" <<PY || tail -5 $out.err
import json,sys
d=json.load(open(sys.argv[1])); e=d.get("e2e") or {}
print("single-process w2 N=$N value", round(d["value"]), "e2e", round(e.get("value",0)), "parity", d.get("parity_ok"), "wall/step incl stitch", d["config"].get("host_wall_ms_per_step_incl_stitch"))
PY
fi
