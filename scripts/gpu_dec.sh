#!/bin/bash
# decoder work loop: parity tests of the decoder paths, then the configs[3] bench (32 boxes / image) with its per-class breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_vitb.py -q -m gpu -s -x -k "decoder or thirty_two" > gpurun_out/dec_tests.log 2>&1; echo "tests exit $?"
grep -h "rel-L2\|IoU\|passed\|failed\|Error\|error" gpurun_out/dec_tests.log | tail -12
timeout 600 python bench.py --workload b32 --steps ${STEPS:-6} --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dec_b32.json 2> gpurun_out/dec_b32.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/dec_b32.json"))
b=d["breakdown"]
print("b32 value %.1f img/s e2e %.1f | %s" % (d["value"], d["e2e"]["value"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and (k.startswith("dec") or k.startswith("post")))))
PY
if [ "${ATTN:-0}" = "1" ]; then timeout 120 python scripts/attn_micro.py "current" 2>&1 | tail -1; fi
