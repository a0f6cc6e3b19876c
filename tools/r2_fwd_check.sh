#!/usr/bin/env bash
# 1 GPU: forward-projector tests + per-kernel bench of the forwards.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forward or session or projector" 2>&1 | tail -4
timeout 900 python tools/kernel_bench.py --skip runs0,runs1,session,scatter,padj > gpurun_out/kb_fwd.json 2> gpurun_out/kb_fwd.err; echo "rc=$?"; tail -3 gpurun_out/kb_fwd.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/kb_fwd.json") if l.startswith("{")][-1])
    for k, v in d.items():
        if isinstance(v, dict) and "ms" in v:
            print("%-34s %8.4f ms  frac %s" % (k, v["ms"], round(v.get("frac", 0), 3)))
except Exception as e:
    print("no line:", e)
PY
