#!/usr/bin/env bash
# Round-2 first 1-GPU call: launch the kernels that were written after round 1's GPU budget was spent
# (tests marked `unrun` in tests/test_gpu_parity.py), each under `timeout`.
#   gpurun --timeout 1200 -- 'bash tools/validate_unrun.sh'
# When this passes, drop the `unrun` marker from those tests (and, if the prepared forward is faster,
# make it bench.py's default --forward).
set -u
mkdir -p gpurun_out
echo "== validated suite"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== not-yet-run kernels"
IONO_TEST_UNRUN=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "gaussian or forward_projector or run_compressed or large_axis" 2>&1 | tail -15
for fwd in sweep prepared runs; do
  echo "== bench variant $fwd (runs = prepared forward + run-compressed binned adjoint)"
  if [ $fwd = runs ]; then export IONO_BP_RUNS=1; f=prepared; else f=$fwd; fi
  timeout 240 python bench.py --forward $f --no-e2e --no-cpu-baseline --steps 30 --warmup 5 \
      > gpurun_out/unrun_bench_$fwd.json 2> gpurun_out/unrun_bench_$fwd.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/unrun_bench_$fwd.json"))
    print({k: round(v["ms"], 3) for k, v in d["kernels"].items()}, "ms/step", round(d["ms_per_step"], 3))
except Exception as e:
    print("no bench line:", e)
PY
done
# ncu only after the plain runs above exited 0 (never under a multi-rank launch)
if [ -s gpurun_out/unrun_bench_runs.json ]; then
  echo "== launch list of the bench with the prepared operators"
  IONO_BP_RUNS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_prepared.csv python bench.py --forward prepared --no-e2e --no-cpu-baseline \
      --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
  echo "rc=$?"
  echo "== ncu --set full of the two apply kernels"
  NT=100 IONO_BP_RUNS=1 timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:"prepared_forward|backproject_wruns" -c 4 -f -o gpurun_out/prof_prepared \
      python tools/profile_prepared.py > gpurun_out/ncu_prepared.log 2>&1
  echo "rc=$?"
fi
