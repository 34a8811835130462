#!/bin/bash
for d in 0 1 2 4 8 3 7 15; do
  YSI_ATTN_DBG=$d timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/abl_$d.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/abl_$d.json'));b=d['breakdown'];print('dbg $d', round(b['attn_global']['ms_per_step'],3), round(b['attn_window']['ms_per_step'],3))"
done
