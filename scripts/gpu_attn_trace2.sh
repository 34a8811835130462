#!/bin/bash
# phase-boundary clock traces of one global and one windowed attention CTA (build with -DYSI_ATTN_TRACE on the box)
YSI_NVCC_DEFINES="-DYSI_ATTN_TRACE ${EXTRA_DEFINES}" python -m yolo_sam_inference_b200.build --force --quiet --precision=fp16 > /dev/null 2>&1 || { echo build failed; exit 1; }
timeout 100 python scripts/attn_case.py global 64 2>/dev/null | grep "^TR" | tail -27
timeout 100 python scripts/attn_case.py window 64 2>/dev/null | grep "^TW" | tail -19
