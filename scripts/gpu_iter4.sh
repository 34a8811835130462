#!/bin/bash
# decoder parity tests, configs[3] bench, then a full ncu capture of the decoder's per-box GEMM kernels of one configs[3] batch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_e2e.py "tests/test_gpu_vitb.py::test_vit_b_thirty_two_boxes_end_to_end" tests/test_gpu_encoder.py -q -m gpu -x -s > gpurun_out/iter4_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter4_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error" gpurun_out/iter4_tests.log | tail -8
timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter4_b32.json 2> gpurun_out/iter4_b32.err; echo "bench b32 exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/iter4_b32.json"))
b = d["breakdown"]
print("b32 value %.1f img/s e2e %.1f clk %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
PY
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"EpiKeysLN|EpiConvT|gemm_op16_kernel<128" -c 8 \
  -o gpurun_out/r02_ncu_dec3 -f python scripts/ncu_batch.py 32 > gpurun_out/ncu_dec3.log 2>&1
echo "full capture b32 exit $?"
