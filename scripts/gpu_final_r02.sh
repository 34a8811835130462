#!/bin/bash
# Round-2 evidence run on the committed code: whole GPU suite (both operand builds through the parametrized fixtures), smoke,
# the default bench line (driver's command), the reference arm, then ncu launch lists of one configs[1] / configs[3] batch and a
# full capture of the top encoder kernels. Every step has its own timeout; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/final_summary.txt
timeout 1700 python -m pytest tests/ -x -q -m gpu -s --durations=10 > gpurun_out/suite.log 2>&1; echo "suite exit $?" >> gpurun_out/final_summary.txt
grep -h "IoU\|rel-L2\|passed\|failed\|Error\|error" gpurun_out/suite.log | tail -n 40 >> gpurun_out/final_summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/final_summary.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench reference exit $?" >> gpurun_out/final_summary.txt
for b in 1 32; do
  timeout 300 python scripts/ncu_batch.py $b > gpurun_out/ncu_plain_b$b.log 2>&1 || { echo "plain run b$b failed" >> gpurun_out/final_summary.txt; continue; }
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/ncu_launches_b$b.csv python scripts/ncu_batch.py $b > gpurun_out/ncu_launches_b$b.log 2>&1
  echo "launch list b$b exit $?" >> gpurun_out/final_summary.txt
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'gemm2_op16_kernel|encoder_attention_kernel' -c 14 \
  -o gpurun_out/r02_ncu_top -f python scripts/ncu_batch.py 1 > gpurun_out/ncu_top.log 2>&1
echo "full capture b1 exit $?" >> gpurun_out/final_summary.txt
cat gpurun_out/final_summary.txt
