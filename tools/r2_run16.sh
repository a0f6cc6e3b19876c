#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/profile_padj.py; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prepared_adjoint -c 1 -f -o gpurun_out/r02c_padj \
    python tools/profile_padj.py > gpurun_out/ncu_padj.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_padj.log
