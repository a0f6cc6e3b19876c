#!/usr/bin/env bash
# Round-2 multi-GPU call: 2-rank parity test, then strong-scaling bench lines (peer and NCCL reducers).
# usage: gpurun --gpus N -- 'bash tools/r2_run2_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
echo "== 2-rank parity"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
for red in peer nccl; do
  echo "== bench N=$N reducer=$red"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 50 --warmup 5 --reducer $red > gpurun_out/bench_n${N}_$red.json 2> gpurun_out/bench_n${N}_$red.err
  echo "rc=$?"; tail -4 gpurun_out/bench_n${N}_$red.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n${N}_$red.json") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "scaling", d["scaling"])
    print({k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    print("e2e", d["e2e"] and d["e2e"]["ms_per_step"], "verify", d["verify"])
except Exception as e:
    print("no bench line:", e)
PY
done
