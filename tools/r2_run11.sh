#!/usr/bin/env bash
# 1 GPU: parity suite, then tools/r2_run10.sh (bench + 1/8 shard emulation).
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash tools/r2_run10.sh
