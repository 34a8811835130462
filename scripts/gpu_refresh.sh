#!/bin/bash
# quick regression + bench refresh (no ncu)
mkdir -p gpurun_out
: > gpurun_out/refresh_summary.txt
for t in encoder e2e postprocess; do
  timeout 300 python -m pytest tests/test_gpu_$t.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/refresh_summary.txt; tail -n 1 gpurun_out/test_$t.log >> gpurun_out/refresh_summary.txt
done
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/refresh_summary.txt
YSI_PRECISION=bf16 timeout 200 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?" >> gpurun_out/refresh_summary.txt
YSI_BENCH_BOXES=32 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench b32 exit $?" >> gpurun_out/refresh_summary.txt
YSI_BENCH_MODEL=vit_h timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vith.json 2> gpurun_out/bench_vith.err; echo "bench vit_h exit $?" >> gpurun_out/refresh_summary.txt
cat gpurun_out/refresh_summary.txt
for f in bench bench_bf16 bench_b32 bench_vith; do python -c "
import json
d=json.load(open('gpurun_out/$f.json'));print('$f', d['dtype'], round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))"; done
