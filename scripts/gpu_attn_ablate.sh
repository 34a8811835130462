#!/bin/bash
# timing-only ablations of the attention softmax loop (WRONG results by construction): which instruction group bounds it?
mkdir -p gpurun_out
for V in ${1:-0 1 2 4 8 16 3 27}; do
  YSI_NVCC_DEFINES="-DYSI_ATTN_ABLATE=$V ${EXTRA_DEFINES}" python -m yolo_sam_inference_b200.build --force --quiet --precision=fp16 > gpurun_out/abl_build_$V.log 2>&1 || { echo "build $V failed"; tail -3 gpurun_out/abl_build_$V.log; continue; }
  timeout 120 python scripts/attn_micro.py "ablate=$V ${EXTRA_DEFINES}" 2>&1 | tail -1
done
