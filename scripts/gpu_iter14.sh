#!/bin/bash
# stable LayerNorm statistics: encoder / ViT-B / ViT-H parity + reproducibility + switch agreement, one configs[1] bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_vitb.py tests/test_gpu_switches.py -q -m gpu -x -s > gpurun_out/iter14_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter14_tests.log)"
grep -h "rel-L2\|IoU\|Error\|error\|FAILED\|assert" gpurun_out/iter14_tests.log | grep -v bf16 | tail -8
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter14_b1.json 2> gpurun_out/iter14_b1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/iter14_b1.json"))
b = d["breakdown"]
print("b1 value %.1f img/s e2e %.1f clk %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and not k.startswith(("dec", "post")))))
PY
