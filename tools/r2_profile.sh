#!/usr/bin/env bash
# Round-2 profiles (1 GPU): launch lists of the bench command and of a rank's share of an 8-way split,
# then ONE ncu --set full capture of the step's kernels.  The plain runs come first (never profile a failing program).
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_prof.json 2> gpurun_out/bench_r2_prof.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/ncu_l1.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_shard8.csv \
    python bench.py --steps 2 --warmup 3 --emulate-shard 8 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo "rc=$?"
NT=100 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"ray_sweep|prepared_forward|backproject_w|backproject_combine|residual|quads|ne_rows" -c 16 -f -o gpurun_out/r02_prof_step \
    python tools/profile_r2.py > gpurun_out/ncu_r2_step.log 2>&1; echo "rc=$?"
SCATTER=1 NT=100 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"ray_sweep_kernel<1" -c 1 -f -o gpurun_out/r02_prof_scatter \
    python tools/profile_r2.py > gpurun_out/ncu_r2_scatter.log 2>&1; echo "rc=$?"
