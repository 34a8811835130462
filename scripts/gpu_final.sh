#!/bin/bash
# Round-end evidence run: full GPU parity suite (default fp16 operands; bf16 subset), smoke, benches, ncu launch list
# and full-set captures of the top kernels. Every step has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
: > gpurun_out/final_summary.txt
for t in gemm attention postprocess metrics encoder decoder e2e vitb; do
  timeout 600 python -m pytest tests/test_gpu_$t.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "fp16 $t exit $?" >> gpurun_out/final_summary.txt
  grep -h "IoU\|rel-L2\|passed\|failed\|rror" gpurun_out/test_$t.log | tail -n 6 >> gpurun_out/final_summary.txt
done
for t in gemm attention decoder encoder e2e; do
  YSI_PRECISION=bf16 timeout 300 python -m pytest tests/test_gpu_$t.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_${t}_bf16.log 2>&1
  echo "bf16 $t exit $?" >> gpurun_out/final_summary.txt
  grep -h "IoU\|passed\|failed\|rror" gpurun_out/test_${t}_bf16.log | tail -n 3 >> gpurun_out/final_summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_summary.txt
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/final_summary.txt
YSI_PRECISION=bf16 timeout 200 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?" >> gpurun_out/final_summary.txt
YSI_BENCH_BOXES=32 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench b32 exit $?" >> gpurun_out/final_summary.txt
YSI_BENCH_MODEL=vit_h timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vith.json 2> gpurun_out/bench_vith.err; echo "bench vit_h exit $?" >> gpurun_out/final_summary.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench reference exit $?" >> gpurun_out/final_summary.txt
# ncu: launch list of the same command, then full-set captures (after the plain runs above exited)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 330 --csv --log-file gpurun_out/ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?" >> gpurun_out/final_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm2_op16_kernel|encoder_attention_kernel" -s 12 -c 12 \
  -o gpurun_out/ncu_top -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_top.log 2>&1; echo "ncu top exit $?" >> gpurun_out/final_summary.txt
YSI_BENCH_BOXES=32 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'upsample_stats_fast|contour_hull_disk|EpiConvT|tok_gemm|t2i_attention' -s 40 -c 8 \
  -o gpurun_out/ncu_post -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_post.log 2>&1; echo "ncu post exit $?" >> gpurun_out/final_summary.txt
cat gpurun_out/final_summary.txt
