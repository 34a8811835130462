#!/bin/bash
# Runs every GPU parity test file in its own process (a hang or fault in one must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/summary.txt
for t in ${TESTS:-gemm attention postprocess metrics encoder decoder e2e}; do
  timeout ${TEST_TIMEOUT:-420} python -m pytest tests/test_gpu_$t.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$t.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
