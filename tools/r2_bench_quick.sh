#!/usr/bin/env bash
# 1 GPU: bench (no cpu baseline) + 1/8 shard emulation; prints step time against the sum of the component times.
set -u
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    ks = {k: round(v["ms"], 4) for k, v in d["kernels"].items()}
    tot = sum(v for k, v in ks.items() if k != "cast_rays")
    print("ms/step", round(d["ms_per_step"], 4), "sum of components", round(tot, 4), "gap", round(d["ms_per_step"] - tot, 4), "pass frac", round(d["pass_frac_of_hbm_roofline"], 4))
    print(ks)
    print("e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3), "verify", d["verify"] and d["verify"]["ok"], "launches", d["gpu_launches"])
except Exception as e:
    print("no bench line:", e)
PY
}
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ${EXTRA:-} > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2e.err
show gpurun_out/bench_r2e.json
echo "== emulate shard 1/8"
timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard 8 --no-cpu-baseline --no-e2e ${EXTRA:-} > gpurun_out/bench_emul8.json 2> gpurun_out/bench_emul8.err; echo "rc=$?"; tail -3 gpurun_out/bench_emul8.err
show gpurun_out/bench_emul8.json
