#!/bin/bash
# A/B of the attention kernel's threads-per-row (YSI_ATTN_SPLIT) on one box: rebuild, bench ViT-B and ViT-H, alternate
for s in 1 2 1 2; do
  YSI_NVCC_DEFINES=-DYSI_ATTN_SPLIT=$s timeout 200 python yolo_sam_inference_b200/build.py --force > /dev/null 2>&1
  timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_b.json 2>/dev/null
  YSI_BENCH_MODEL=vit_h timeout 150 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_h.json 2>/dev/null
  python -c "
import json
for m in 'bh':
    d=json.load(open('gpurun_out/ab_%s.json'%m));b=d['breakdown']
    print('split $s', m, round(d['value'],1), d['clocks']['sm_mhz'], 'global', round(b['attn_global']['ms_per_step'],3), 'window', round(b['attn_window']['ms_per_step'],3), 'fc1', round(b['gemm_fc1']['ms_per_step'],3))"
done
