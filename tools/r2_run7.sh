#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== kernel bench (stateless adjoint variants)"
timeout 600 python tools/kernel_bench.py --skip sweep,prepared,runs0,session > gpurun_out/kb_r2c.json 2> gpurun_out/kb_r2c.err; echo "rc=$?"; tail -3 gpurun_out/kb_r2c.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/kb_r2c.json"))
    print({k: (round(v["ms"], 3), round(v.get("frac", 0), 3)) if isinstance(v, dict) and "ms" in v else v
           for k, v in d.items() if k not in ("env",)})
except Exception as e:
    print("failed:", e)
PY
