#!/bin/bash
# final verification of the committed code: whole GPU suite (fp16 default), bf16 subset, smoke, default bench
mkdir -p gpurun_out
: > gpurun_out/verify_summary.txt
timeout 800 python -m pytest tests/ -x -q -m gpu -s > gpurun_out/verify_fp16.log 2>&1; echo "fp16 suite exit $?" >> gpurun_out/verify_summary.txt
grep -h "IoU\|embeddings\|decoder low-res\|passed\|failed" gpurun_out/verify_fp16.log | tail -n 14 >> gpurun_out/verify_summary.txt
YSI_PRECISION=bf16 timeout 600 python -m pytest tests/ -x -q -m gpu -s -k "not vit_h and not pipeline" > gpurun_out/verify_bf16.log 2>&1; echo "bf16 suite exit $?" >> gpurun_out/verify_summary.txt
grep -h "IoU\|passed\|failed" gpurun_out/verify_bf16.log | tail -n 6 >> gpurun_out/verify_summary.txt
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/verify_summary.txt
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/verify_summary.txt
YSI_BENCH_MODEL=vit_h timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vith.json 2> gpurun_out/bench_vith.err; echo "bench vit_h exit $?" >> gpurun_out/verify_summary.txt
cat gpurun_out/verify_summary.txt
python -c "
import json
for f in ('bench','bench_vith'):
    d=json.load(open('gpurun_out/%s.json'%f));print(f, d['dtype'], round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))"
