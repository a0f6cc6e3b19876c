#!/usr/bin/env bash
# 1 GPU: parity (new optimiser / host-session tests), bench with the e2e variants, inversion (configs[3]).
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "rc=$?"; tail -3 gpurun_out/bench_r2e.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r2e.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 4), "pass frac", round(d["pass_frac_of_hbm_roofline"], 4))
    for k in ("e2e", "e2e_active_voxels", "e2e_cold"):
        e = d.get(k)
        print(k, e and (round(e["ms_per_step"], 3), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]))
except Exception as e:
    print("no bench line:", e)
PY
echo "== inversion C2-size grid"
timeout 600 python tools/bench_inversion.py --grid 256 256 128 --iters 50 > gpurun_out/inv_c2_r2.json 2> gpurun_out/inv_c2_r2.err; echo "rc=$?"; tail -2 gpurun_out/inv_c2_r2.err; cat gpurun_out/inv_c2_r2.json
echo "== inversion C4 (512x512x256, Ns=256)"
timeout 900 python tools/bench_inversion.py --iters 50 > gpurun_out/inv_c4_r2.json 2> gpurun_out/inv_c4_r2.err; echo "rc=$?"; tail -2 gpurun_out/inv_c4_r2.err; cat gpurun_out/inv_c4_r2.json
