#!/bin/bash
# A/B bench of an env knob: usage  gpu_ab.sh VAR v1 v2 ...
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps ${STEPS:-16} --warmup 3 --no-cpu-baseline > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  echo "$VAR=$v exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${VAR}_$v.json"))
print(" value %.1f e2e %.1f gemm %.0f TF" % (d["value"], d["e2e"]["value"], d["roofline"]["achieved"]), {k: round(v["ms_per_step"],3) for k,v in d["breakdown"].items() if isinstance(v, dict)})
PY
done
