#!/bin/bash
# A/B of attention-kernel build options on one box: correctness (attention tests + ViT-B 8 images) for the default build,
# then the b1 bench (no extras) per variant.   usage: scripts/gpu_ab_attn.sh "<defines variant 1>|<defines variant 2>|..."
mkdir -p gpurun_out
python -m yolo_sam_inference_b200.build --quiet > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_vitb.py tests/test_gpu_encoder.py -q -m gpu -s -k "attention or config0 or reproducible" > gpurun_out/ab_tests.log 2>&1; echo "tests exit $?"
grep -h "IoU\|passed\|failed\|Error" gpurun_out/ab_tests.log | tail -8
IFS='|' read -ra VARS <<< "${1:-default}"
k=0
for V in "${VARS[@]}"; do
  k=$((k+1))
  DEF="$V"; [ "$V" = "default" ] && DEF=""
  YSI_NVCC_DEFINES="$DEF" python -m yolo_sam_inference_b200.build --force --quiet --precision=fp16 > gpurun_out/ab_build_$k.log 2>&1 || { echo "build $V failed"; tail -5 gpurun_out/ab_build_$k.log; continue; }
  for G in ${GRAPHS:-1}; do
  YSI_GRAPH=$G timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ab_$k.json 2> gpurun_out/ab_$k.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$k.json"))
b=d["breakdown"]
print("[$V] graph=$G value %.1f e2e %.1f clocks %s | attn_global %.3f attn_window %.3f ln %.3f enc_tflops %.1f" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], b["attn_global"]["ms_per_batch"], b["attn_window"]["ms_per_batch"], b["layernorm"]["ms_per_batch"], b["_encoder_alg_tflops"]))
PY
  done
done
