#!/bin/bash
# phase-boundary clock trace of one global-attention CTA: build the fp16 library with -DYSI_ATTN_TRACE first:
#   YSI_NVCC_DEFINES=-DYSI_ATTN_TRACE python yolo_sam_inference_b200/build.py --force   (and rebuild without it afterwards)
timeout 100 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep "^TR" | head -27
