#!/usr/bin/env bash
# 1 GPU: parity, bench, shard emulation.
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
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2d.err
show gpurun_out/bench_r2d.json
for n in 8; do
  echo "== emulate shard 1/$n"
  timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard $n --no-cpu-baseline > gpurun_out/bench_emul$n.json 2> gpurun_out/bench_emul$n.err; echo "rc=$?"; tail -3 gpurun_out/bench_emul$n.err
  show gpurun_out/bench_emul$n.json
done
