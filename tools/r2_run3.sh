#!/usr/bin/env bash
# Round-2 GPU call 3 (1 GPU): bench line, the per-rank step of an 8-way split emulated on one GPU, launch lists.
set -u
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], "pass frac", round(d["pass_frac_of_hbm_roofline"], 4))
    print({k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    print("e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3), "verify", d["verify"] and d["verify"]["ok"], "launches", d["gpu_launches"])
except Exception as e:
    print("no bench line:", e)
PY
}
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2c.err
show gpurun_out/bench_r2c.json
for n in 2 8; do
  echo "== emulate shard 1/$n"
  timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard $n --no-cpu-baseline > gpurun_out/bench_emul$n.json 2> gpurun_out/bench_emul$n.err; echo "rc=$?"; tail -3 gpurun_out/bench_emul$n.err
  show gpurun_out/bench_emul$n.json
done
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2c.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/ncu_l1.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2c_emul8.csv \
    python bench.py --steps 2 --warmup 1 --emulate-shard 8 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo "rc=$?"
