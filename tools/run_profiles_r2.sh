#!/bin/bash
# Round-2 single-GPU evidence (run under gpurun, one GPU): bench lines, launch list and ncu captures of the SAME commands,
# each ncu pass only after the plain command has exited 0.  Raw outputs land in gpurun_out/; tools/summarise_r2.py turns
# them into the committed summaries under profiles/.
set -x
O=gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
python bench.py > $O/r2_bench_headline.json 2> $O/r2_bench_headline.err; tail -2 $O/r2_bench_headline.err
python bench.py --mutual 0 --no-cpu-baseline --no-int8-peak > $O/r2_bench_mutual0.json 2> $O/r2_bench_mutual0.err
# launch list of the headline step (this repo's kernels only: the synthetic-data generator's torch kernels are not part of a step)
K='match_pairs_kernel|select_candidates|emit_matches|scan_counts|gather_matches|pull_plan|tile_min|pack_|init_pad'
A="--no-cpu-baseline --no-e2e --no-int8-peak --parity-pairs 0"
python bench.py $A --steps 2 --warmup 1 > $O/plain_r2_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $O/r2_launches_headline.csv python bench.py $A --steps 2 --warmup 1 > $O/ncu_launches_r2.log 2>&1
# full capture of the headline workload's forward launch: first batch (2,048 pairs) of the timed step (launch 6 = after the
# warm-up step's 3 forward + 3 gated twin launches)
python bench.py $A --steps 1 --warmup 1 > $O/plain_r2_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_pairs_kernel -s 6 -c 1 -o $O/r2_match_headline -f python bench.py $A --steps 1 --warmup 1 > $O/ncu_full_r2.log 2>&1
tail -2 $O/ncu_full_r2.log
# config #3 on one GPU: bench line + one forward launch of it under ncu
python bench.py --workload 3 --steps 3 --warmup 2 --e2e-steps 3 --no-cpu-baseline --no-int8-peak > $O/r2_scale_w3_1gpu.json 2> $O/r2_scale_w3_1gpu.err
python bench.py --workload 3 $A --steps 1 --warmup 1 > $O/plain_r2_w3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_pairs_kernel -s 8 -c 1 -o $O/r2_match_w3 -f python bench.py --workload 3 $A --steps 1 --warmup 1 > $O/ncu_full_r2_w3.log 2>&1
tail -2 $O/ncu_full_r2_w3.log
python bench.py --workload 2 --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > $O/r2_scale_w2_1gpu.json 2> $O/r2_scale_w2_1gpu.err
