#!/bin/bash
# parity tests + short bench for both operand encodings
mkdir -p gpurun_out
: > gpurun_out/prec_summary.txt
for prec in fp16 bf16; do
  export YSI_PRECISION=$prec
  for t in ${TESTS:-gemm attention decoder encoder e2e vitb}; do
    sel=""; [ "$t" = vitb ] && sel="-k not_vit_h"
    timeout 600 python -m pytest tests/test_gpu_$t.py -q -m gpu -s -x --no-header -p no:cacheprovider ${sel:+-k "not vit_h"} > gpurun_out/test_${t}_$prec.log 2>&1
    echo "$prec $t exit $?" >> gpurun_out/prec_summary.txt
    grep -h "IoU\|rel-L2\|passed\|failed\|Error" gpurun_out/test_${t}_$prec.log | tail -n 8 >> gpurun_out/prec_summary.txt
  done
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
  echo "$prec bench exit $?" >> gpurun_out/prec_summary.txt
  python -c "import json;d=json.load(open('gpurun_out/bench_$prec.json'));print(d['dtype'],d['value'],d['e2e']['value'],d['roofline']['achieved'])" >> gpurun_out/prec_summary.txt 2>&1
done
cat gpurun_out/prec_summary.txt
