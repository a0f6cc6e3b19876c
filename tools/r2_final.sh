#!/usr/bin/env bash
# Final 1-GPU record of round 2: the bench line (with the CPU baseline), launch lists, ONE ncu --set full capture of the
# step's kernels, per-kernel bench, the inversion benchmark.  Plain runs first (never profile a failing program).
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench (default flags)"
timeout 1500 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "rc=$?"; tail -2 gpurun_out/final_bench_n1.err
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "rc=$?"
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/ncu_l1.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches_shard8.csv \
    python bench.py --steps 2 --warmup 3 --emulate-shard 8 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo "rc=$?"
echo "== ncu full"
NT=100 timeout 300 python tools/profile_r2.py; echo "plain rc=$?"
# (one pass, 12 kernels: the report must stay well under gpurun's 64 MiB return limit)
PASSES=1 SCATTER=1 NT=100 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"ray_sweep|prepared_forward|prepared_adjoint|finish_|backproject_w|backproject_combine|residual|quads|ne_rows|adjoint_runs" -c 12 -f -o gpurun_out/final_prof_step \
    python tools/profile_r2.py > gpurun_out/ncu_final_step.log 2>&1; echo "rc=$?"
echo "== kernel bench"
timeout 900 python tools/kernel_bench.py > gpurun_out/final_kernel_bench.json 2> gpurun_out/final_kernel_bench.err; echo "rc=$?"
echo "== inversion"
timeout 900 python tools/bench_inversion.py > gpurun_out/final_inv_c4.json 2> gpurun_out/final_inv_c4.err; echo "rc=$?"
timeout 900 python tools/bench_inversion.py --grid 256 256 128 > gpurun_out/final_inv_c2.json 2> gpurun_out/final_inv_c2.err; echo "rc=$?"
timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard 8 --no-cpu-baseline --no-e2e > gpurun_out/final_bench_emul8.json 2> gpurun_out/final_bench_emul8.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("final_bench_n1", "final_bench_ref", "final_bench_emul8"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "pass_frac_of_hbm_roofline", "gpu_launches", "impl")}, d.get("e2e") and d["e2e"].get("ms_per_step"), d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
