#!/bin/bash
# LayerNorm grid A/B on configs[1] + encoder parity tests, then the launch list (duration + DRAM bytes) of one configs[3] batch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -m gpu -x > gpurun_out/iter3_tests.log 2>&1; echo "tests exit $? $(tail -1 gpurun_out/iter3_tests.log)"
summ() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
b = d["breakdown"]
print("%s value %.1f img/s e2e %.1f clk %s | %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
PY
}
for f in 3 0 4 2; do
  YSI_LN_CTAS_PER_SM=$f timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter3_b1_ln$f.json 2> gpurun_out/iter3_b1_ln$f.err; echo "bench b1 ln_ctas=$f exit $?"
  summ gpurun_out/iter3_b1_ln$f.json "b1 ln_ctas_per_sm=$f"
done
timeout 300 python scripts/ncu_batch.py 32 > gpurun_out/ncu_plain_b32.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_b32.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/ncu_launches_b32.csv python scripts/ncu_batch.py 32 > gpurun_out/ncu_launches_b32.log 2>&1
echo "launch list b32 exit $?"
