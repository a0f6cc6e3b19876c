#!/usr/bin/env bash
# 1 GPU: full parity suite; bench (default adjoint = prepared) + 1/8 emulation; bench with --adjoint binned; C4 L-BFGS both adjoints.
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
bash tools/r2_run10.sh
cp gpurun_out/bench_r2e.json gpurun_out/bench_r2e_prepared.json; cp gpurun_out/bench_emul8.json gpurun_out/bench_emul8_prepared.json
echo "== binned adjoint"
EXTRA="--adjoint binned --no-e2e" bash tools/r2_run10.sh
echo "== e2e with the prepared adjoint"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-verify --e2e-adjoint prepared > gpurun_out/bench_e2e_prep.json 2>gpurun_out/bench_e2e_prep.err; echo rc=$?
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_e2e_prep.json") if l.startswith("{")][-1])
    print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["adjoint"], "active", d["e2e_active_voxels"]["ms_per_step"], d["e2e_active_voxels"]["adjoint"])
except Exception as e:
    print("no line", e)
PY
for adj in prepared binned; do
echo "== C4 L-BFGS $adj"
timeout 900 python tools/bench_inversion.py --adjoint $adj > gpurun_out/inv_c4_$adj.json 2> gpurun_out/inv_c4_$adj.err; echo rc=$?; tail -2 gpurun_out/inv_c4_$adj.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/inv_c4_$adj.json") if l.startswith("{")][-1])
    print({k: d[k] for k in d if k in ("ms_per_iteration_steady", "ms_per_iteration", "forward_ms", "adjoint_ms", "operator_gb", "adjoint", "share_of_time_in_forward_and_adjoint", "S_first", "S_last")})
except Exception as e:
    print("no line", e)
PY
done
