#!/usr/bin/env bash
# 1 GPU: the prepared-adjoint tests, then the per-kernel bench of the adjoints.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adjoint or session" 2>&1 | tail -8
timeout 900 python tools/kernel_bench.py --skip runs0,session,sweep,prepared,scatter > gpurun_out/kb_r2f.json 2> gpurun_out/kb_r2f.err; echo "rc=$?"; tail -3 gpurun_out/kb_r2f.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/kb_r2f.json") if l.startswith("{")][-1])
    for k, v in d.items():
        if isinstance(v, dict) and "ms" in v:
            print("%-34s %8.4f ms  frac %s" % (k, v["ms"], round(v.get("frac", 0), 3)))
        elif not isinstance(v, dict):
            print(k, v)
except Exception as e:
    print("no line:", e)
PY
