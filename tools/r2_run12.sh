#!/usr/bin/env bash
# 1 GPU: parity suite, then ncu --set full of the prepared forward, the binned apply and the run-aggregated scatter adjoint.
set -u
mkdir -p gpurun_out
echo "== parity"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
NT=100 timeout 300 python tools/profile_r2.py; echo "plain rc=$?"
SCATTER=1 NT=100 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"prepared_forward|backproject_wruns|adjoint_runs" -c 5 -f -o gpurun_out/r02b_prof \
    python tools/profile_r2.py > gpurun_out/ncu_r2b.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_r2b.log
