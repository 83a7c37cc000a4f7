set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python bench.py > gpurun_out/r1_bench_mutual1.json 2> gpurun_out/r1_bench_mutual1.err
python bench.py --mutual 0 --no-cpu-baseline > gpurun_out/r1_bench_mutual0.json 2> gpurun_out/r1_bench_mutual0.err
tail -2 gpurun_out/r1_bench_mutual1.err
# launch list of the same command (short: 32 images, 2 steps) and a full capture of the forward + mutual launches
python bench.py --images 32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-int8-peak > gpurun_out/plain_r1c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --images 32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-int8-peak > gpurun_out/ncu_launches_r1c.log 2>&1
python bench.py --images 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-int8-peak > gpurun_out/plain_r1d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:match_pairs -s 2 -c 2 -o gpurun_out/prof_r1c -f python bench.py --images 32 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-int8-peak > gpurun_out/ncu_full_r1c.log 2>&1
tail -2 gpurun_out/ncu_full_r1c.log
