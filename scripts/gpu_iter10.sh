#!/bin/bash
# decoder parity tests + configs[3] bench (work loop for decoder changes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_e2e.py "tests/test_gpu_vitb.py::test_vit_b_thirty_two_boxes_end_to_end" -q -m gpu -x -s > gpurun_out/iter10_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter10_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error" gpurun_out/iter10_tests.log | grep -v bf16 | tail -6
for k in 1 2; do
timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter10_b32_$k.json 2> gpurun_out/iter10_b32_$k.err; echo "bench b32 exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/iter10_b32_$k.json"))
b = d["breakdown"]
print("b32 value %.1f img/s e2e %.1f clk %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and k.startswith(("dec", "post")))))
PY
done
