#!/bin/bash
# bench sweeps: residual-add mode and images per step (device-resident leg only matters here)
mkdir -p gpurun_out
for cfg in "1 8" "2 8" "2 4" "2 16"; do
  set -- $cfg
  YSI_RESIDUAL_MODE=$1 YSI_BENCH_BATCH=$2 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/sweep_r$1_b$2.json 2> gpurun_out/sweep_r$1_b$2.err
  echo "mode $1 batch $2 exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_r$1_b$2.json"))
print(" value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_step"],3) for k,v in d["breakdown"].items() if isinstance(v, dict)})
PY
done
