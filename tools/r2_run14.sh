#!/usr/bin/env bash
# 1 GPU: parity suite, per-kernel bench incl. the prepared adjoint, bench with --adjoint prepared (+ 1/8 emulation).
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== kernel bench"
timeout 900 python tools/kernel_bench.py --skip runs0,session,sweep > gpurun_out/kb_r2e.json 2> gpurun_out/kb_r2e.err; echo "rc=$?"; tail -3 gpurun_out/kb_r2e.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/kb_r2e.json") if l.startswith("{")][-1])
    for k, v in d.items():
        if isinstance(v, dict) and "ms" in v:
            print("%-34s %8.4f ms  frac %s" % (k, v["ms"], round(v.get("frac", 0), 3)))
        elif not isinstance(v, dict):
            print(k, v)
except Exception as e:
    print("no line:", e)
PY
EXTRA="--adjoint prepared" bash tools/r2_run10.sh
