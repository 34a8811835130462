#!/bin/bash
# work loop of this session: parity tests touched by the decoder epilogue fusion and the implicit-GEMM neck, then
# configs[3] (32 boxes / image) with / without the fused out-projection + LayerNorm4, configs[1] with / without the implicit neck
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_e2e.py tests/test_gpu_encoder.py tests/test_gpu_vitb.py -q -m gpu -x -s > gpurun_out/iter2_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter2_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error" gpurun_out/iter2_tests.log | tail -14
summ() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
b = d["breakdown"]
print("%s value %.1f img/s e2e %.1f clk %s | %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
PY
}
for f in 1 0; do
  YSI_DEC_FUSED_LN=$f timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter2_b32_f$f.json 2> gpurun_out/iter2_b32_f$f.err; echo "bench b32 fused=$f exit $?"
  summ gpurun_out/iter2_b32_f$f.json "b32 fused_ln=$f"
done
for f in 1 0; do
  YSI_NECK_IMPLICIT=$f timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter2_b1_n$f.json 2> gpurun_out/iter2_b1_n$f.err; echo "bench b1 implicit=$f exit $?"
  summ gpurun_out/iter2_b1_n$f.json "b1 neck_implicit=$f"
done
