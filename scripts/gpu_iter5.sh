#!/bin/bash
# folded LayerNorm: parity / reproducibility tests, configs[1] with and without it, then ncu of the decoder's GEMM kernels (32 boxes)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_vitb.py tests/test_gpu_e2e.py tests/test_gpu_gemm.py -q -m gpu -x -s > gpurun_out/iter5_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter5_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error\|FAILED\|assert" gpurun_out/iter5_tests.log | tail -16
summ() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
b = d["breakdown"]
print("%s value %.1f img/s e2e %.1f clk %s enc_tflops %.0f | %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], b["_encoder_alg_tflops"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
PY
}
for f in 1 0; do
  YSI_LN_FUSED=$f timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter5_b1_ln$f.json 2> gpurun_out/iter5_b1_ln$f.err; echo "bench b1 ln_fused=$f exit $?"
  summ gpurun_out/iter5_b1_ln$f.json "b1 ln_fused=$f"
done
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"^gemm_op16_kernel" -c 8 \
  -o gpurun_out/r02_ncu_dec3 -f python scripts/ncu_batch.py 32 > gpurun_out/ncu_dec3.log 2>&1
echo "full capture b32 exit $?"; ls -la gpurun_out/r02_ncu_dec3.ncu-rep
fi
