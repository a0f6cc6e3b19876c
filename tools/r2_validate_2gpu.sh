#!/usr/bin/env bash
# 2 GPUs: full parity suite (incl. the 2-rank test) and the N=2 strong-scaling bench line.
set -u
mkdir -p gpurun_out
echo "== parity (all gpu tests, 2 GPUs visible)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
echo "rc=$?"; tail -2 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n2_final.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 4), d["ms_per_step_ranks"], "e2e", d["e2e"]["ms_per_step"], "active", d["e2e_active_voxels"]["ms_per_step"], d["verify"])
    print(d["kernel_ms_ranks"]); print(d["kernels"]["peer_reduce_expand"])
except Exception as e:
    print("no bench line:", e)
PY
