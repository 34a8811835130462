#!/bin/bash
# decoder work loop: parity tests that touch the decoder, then configs[3] (32 boxes / image) with and without the fused
# out-projection + LayerNorm4 epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_e2e.py "tests/test_gpu_vitb.py::test_vit_b_thirty_two_boxes_end_to_end" -q -m gpu -x -s > gpurun_out/dec2_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/dec2_tests.log)"
grep -h "rel-L2\|IoU" gpurun_out/dec2_tests.log | tail -12
for f in 1 0; do
YSI_DEC_FUSED_LN=$f timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dec2_b32_f$f.json 2> gpurun_out/dec2_b32_f$f.err; echo "bench b32 fused=$f exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/dec2_b32_f$f.json"))
b=d["breakdown"]
print("fused=$f b32 value %.1f img/s e2e %.1f | %s" % (d["value"], d["e2e"]["value"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and (k.startswith("dec") or k.startswith("post")))))
PY
done
