#!/bin/bash
# token MLP on the tensor cores: decoder parity (72 boxes exercise it), configs[3] with / without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py "tests/test_gpu_vitb.py::test_vit_b_thirty_two_boxes_end_to_end" -q -m gpu -x -s > gpurun_out/iter11_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter11_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error" gpurun_out/iter11_tests.log | tail -8
for f in 1 0; do
YSI_DEC_MLP_TC=$f timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter11_b32_$f.json 2> gpurun_out/iter11_b32_$f.err; echo "bench b32 mlp_tc=$f exit $?"; tail -c 300 gpurun_out/iter11_b32_$f.err
python - <<PY
import json
d = json.load(open("gpurun_out/iter11_b32_$f.json"))
b = d["breakdown"]
print("mlp_tc=$f b32 value %.1f img/s e2e %.1f clk %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and k.startswith(("dec", "post")))))
PY
done
