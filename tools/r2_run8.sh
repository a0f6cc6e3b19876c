#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== kernel bench"
timeout 600 python tools/kernel_bench.py --skip sweep,prepared,runs0,scatter > gpurun_out/kb_r2d.json 2> gpurun_out/kb_r2d.err; echo "rc=$?"; tail -3 gpurun_out/kb_r2d.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/kb_r2d.json"))
    print({k: (round(v["ms"], 3), round(v.get("frac", 0), 3)) if isinstance(v, dict) and "ms" in v else v
           for k, v in d.items() if k not in ("env",)})
except Exception as e:
    print("failed:", e)
PY
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2f.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2f.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 4), "pass frac", round(d["pass_frac_of_hbm_roofline"], 4), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    print("e2e", round(d["e2e"]["ms_per_step"], 3), "verify", d["verify"])
except Exception as e:
    print("no bench line:", e)
PY
