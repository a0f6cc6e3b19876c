#!/usr/bin/env bash
# Round-2 GPU call 2 (1 GPU): parity, bench line, launch-shape variants.
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "rc=$?"
tail -5 gpurun_out/bench_r2b.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2b.json") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "pass frac", d["pass_frac_of_hbm_roofline"])
    print({k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    print("e2e", d["e2e"] and d["e2e"]["ms_per_step"], "cold", d["e2e_cold"] and d["e2e_cold"]["ms_per_step"], "verify", d["verify"])
    print("launches", d["gpu_launches"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
PY
echo "== kernel bench"
timeout 600 python tools/kernel_bench.py --skip sweep,scatter,runs0 > gpurun_out/kb_r2b.json 2> gpurun_out/kb_r2b.err; echo "rc=$?"
tail -3 gpurun_out/kb_r2b.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/kb_r2b.json"))
    print({k: (round(v["ms"], 3), round(v.get("frac", 0), 3)) if isinstance(v, dict) and "ms" in v else v
           for k, v in d.items() if k not in ("env",)})
except Exception as e:
    print("failed:", e)
PY
