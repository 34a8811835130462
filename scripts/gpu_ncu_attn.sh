#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'encoder_attention_kernel' -s 30 -c 4 \
  -o gpurun_out/ncu_attn -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "exit $?"
