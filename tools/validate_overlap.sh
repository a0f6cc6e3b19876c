#!/usr/bin/env bash
# Round-2 first call (2 GPUs, everything under `timeout` so that a dead-locked collective cannot
# eat the GPU budget again):
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/validate_overlap.sh'
set -u
mkdir -p gpurun_out
echo "== multi-rank NCCL test incl. the overlapped apply"
IONO_TEST_OVERLAP=1 timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for ov in 0 4; do
  echo "== bench N=2 overlap=$ov"
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 2960$ov bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --overlap $ov \
      > gpurun_out/overlap_n2_ov$ov.json 2> gpurun_out/overlap_n2_ov$ov.err
  echo "rc=$?"; tail -c 400 gpurun_out/overlap_n2_ov$ov.json; echo
done
