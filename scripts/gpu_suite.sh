#!/bin/bash
# whole GPU parity suite (both operand builds through the parametrized fixtures), smoke, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/suite_summary.txt
timeout ${SUITE_TIMEOUT:-1700} python -m pytest tests/ -x -q -m gpu -s --durations=15 > gpurun_out/suite.log 2>&1; echo "suite exit $?" >> gpurun_out/suite_summary.txt
grep -h "IoU\|rel-L2\|passed\|failed\|Error\|error" gpurun_out/suite.log | tail -n 40 >> gpurun_out/suite_summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/suite_summary.txt
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/suite_summary.txt
  tail -c 600 gpurun_out/bench.err >> gpurun_out/suite_summary.txt
fi
cat gpurun_out/suite_summary.txt
