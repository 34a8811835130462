#!/bin/bash
# work loop: (optional) attention clock trace, then the default build: quick parity tests, attention micro-benchmark, configs[3] bench
mkdir -p gpurun_out
if [ "${TRACE:-0}" = "1" ]; then bash scripts/gpu_attn_trace2.sh; fi
python -m yolo_sam_inference_b200.build --force --quiet > /dev/null 2>&1 || { echo build failed; exit 1; }
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_attention.py tests/test_gpu_postprocess.py tests/test_gpu_metrics.py tests/test_gpu_decoder.py} -q -m gpu -x > gpurun_out/iter_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter_tests.log)"
timeout 120 python scripts/attn_micro.py "current" 2>&1 | tail -1
if [ "${B32:-1}" = "1" ]; then
timeout 600 python bench.py --workload b32 --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dec_b32.json 2> gpurun_out/dec_b32.err; echo "bench b32 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/dec_b32.json"))
b=d["breakdown"]
print("b32 value %.1f img/s e2e %.1f | %s" % (d["value"], d["e2e"]["value"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and (k.startswith("dec") or k.startswith("post")))))
PY
fi
if [ "${B1:-0}" = "1" ]; then
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter_b1.json 2> gpurun_out/iter_b1.err; echo "bench b1 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/iter_b1.json"))
b=d["breakdown"]
print("b1 value %.1f img/s e2e %.1f clocks %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
print("enc tflops", b["_encoder_alg_tflops"])
PY
fi
