#!/usr/bin/env bash
# compute-sanitizer (memcheck, racecheck) over small invocations of every hot kernel, if the tool is usable on the box.
set -u
mkdir -p gpurun_out
which compute-sanitizer || ls /usr/local/cuda/bin/compute-sanitizer
SEL="test_device_session_matches_separate_calls or test_fused_residual_kernel or test_forward_quads_bit_identical or test_binned_backprojector_run_compressed or test_binned_backprojector_chunked_apply or test_optimiser_vector_kernels or test_forward_projector_edges"
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|passed|failed|Error|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
