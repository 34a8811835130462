#!/bin/bash
# patch-embed epilogue through EpiResidLN: parity tests, then configs[1] / ViT-H with and without the folded LayerNorm1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_vitb.py tests/test_gpu_e2e.py -q -m gpu -x -s > gpurun_out/iter7_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter7_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error\|FAILED\|assert" gpurun_out/iter7_tests.log | grep -v bf16 | tail -8
summ() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
b = d["breakdown"]
print("%s value %.1f img/s e2e %.1f clk %s enc_tflops %.0f | %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], b["_encoder_alg_tflops"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and not k.startswith(("dec", "post")))))
PY
}
for f in 1 0; do
  YSI_LN_FUSED=$f timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter7_b1_ln$f.json 2> gpurun_out/iter7_b1_ln$f.err; echo "bench b1 ln_fused=$f exit $?"
  summ gpurun_out/iter7_b1_ln$f.json "b1 ln_fused=$f"
done
for f in 1 0; do
  YSI_LN_FUSED=$f timeout 600 python bench.py --workload vit_h --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter7_vh_ln$f.json 2> gpurun_out/iter7_vh_ln$f.err; echo "bench vit_h ln_fused=$f exit $?"
  summ gpurun_out/iter7_vh_ln$f.json "vit_h ln_fused=$f"
done
