#!/bin/bash
# A/B of attention build options with the kernel-only micro-benchmark: per variant, build the fp16 library, run the attention
# parity tests against it (-k fp16) and time the four attention shapes.   usage: scripts/gpu_ab_micro.sh "<defines 1>|<defines 2>|..."
mkdir -p gpurun_out
IFS='|' read -ra VARS <<< "${1:-default}"
k=0
for V in "${VARS[@]}"; do
  k=$((k+1))
  DEF="$V"; [ "$V" = "default" ] && DEF=""
  YSI_NVCC_DEFINES="$DEF" python -m yolo_sam_inference_b200.build --force --quiet --precision=fp16 > gpurun_out/abm_build_$k.log 2>&1 || { echo "build $V failed"; tail -5 gpurun_out/abm_build_$k.log; continue; }
  timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -k fp16 -x > gpurun_out/abm_test_$k.log 2>&1; echo "[$V] tests exit $? $(tail -1 gpurun_out/abm_test_$k.log)"
  for rep in 1 2; do timeout 120 python scripts/attn_micro.py "$V" 2>&1 | tail -1; done
done
