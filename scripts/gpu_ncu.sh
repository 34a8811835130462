#!/bin/bash
# ncu evidence for profiles/: launch list of one short bench run + full-set captures of the top kernels.
# Only run after the plain bench command exited 0 (numbers printed under ncu are never bench values).
mkdir -p gpurun_out
export YSI_BENCH_BATCH=${YSI_BENCH_BATCH:-8}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 330 --csv --log-file gpurun_out/ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel|encoder_attention_kernel' -s 60 -c 10 \
  -o gpurun_out/ncu_top -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_top.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/*.ncu-rep
