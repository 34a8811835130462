#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16_kernel' -s 3 -c 1 \
  -o gpurun_out/ncu_gemm_qkv -f python scripts/gemm_micro.py qkv 2 0 > gpurun_out/ncu_gemm_qkv.log 2>&1
echo "exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16_kernel' -s 3 -c 1 \
  -o gpurun_out/ncu_gemm_qkv_drain -f python scripts/gemm_micro.py qkv 1 2 > gpurun_out/ncu_gemm_qkv_drain.log 2>&1
echo "exit $?"
