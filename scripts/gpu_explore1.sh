#!/bin/bash
# exploration: 2048^2 16-bit device-side breakdown, batch-size sweep of the default workload
mkdir -p gpurun_out
timeout 300 python scripts/prof_2048.py 2048 1 > gpurun_out/prof_2048.txt 2>&1; echo "prof2048 exit $?"
cat gpurun_out/prof_2048.txt | tail -25
for b in 4 16; do
  YSI_BENCH_BATCH=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_batch$b.json 2> gpurun_out/bench_batch$b.err; echo "batch $b exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_batch$b.json"))
b=d["breakdown"]
print("batch $b value %.1f e2e %.1f clocks %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict))))
PY
done
