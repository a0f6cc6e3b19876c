#!/usr/bin/env bash
# Round-2 GPU call 1: parity suite with the new kernels, per-kernel timings, ncu of the hot kernels.
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== kernel bench"
timeout 600 python tools/kernel_bench.py > gpurun_out/kb_default.json 2> gpurun_out/kb_default.err; echo "rc=$?"
IONO_LIB=tools/variants/libionob200_shfl.so timeout 300 python tools/kernel_bench.py --skip prepared,runs0,runs1,scatter,session \
    > gpurun_out/kb_shfl.json 2> gpurun_out/kb_shfl.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("kb_default", "kb_shfl"):
    try:
        d = json.load(open("gpurun_out/%s.json" % n))
        print(n, {k: (round(v["ms"], 3), round(v.get("frac", 0), 3)) if isinstance(v, dict) and "ms" in v else v
                  for k, v in d.items() if k not in ("env",)})
    except Exception as e:
        print(n, "failed:", e)
PY
echo "== ncu"
NT=100 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"ray_sweep|prepared_forward|backproject_w|residual|quads" -c 12 -f -o gpurun_out/prof_r2a \
    python tools/profile_r2.py > gpurun_out/ncu_r2a.log 2>&1
echo "rc=$?"
