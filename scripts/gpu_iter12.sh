#!/bin/bash
# tensor-core token path at 1 box per image (56 token rows per launch)? configs[1] with the row threshold at 1 and at the default
mkdir -p gpurun_out
for f in 1 448 1 448; do
YSI_DEC_TC_MIN_ROWS=$f timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter12_b1_$f.json 2> gpurun_out/iter12_b1_$f.err; echo "bench b1 min_rows=$f exit $?"; tail -c 300 gpurun_out/iter12_b1_$f.err
python - <<PY
import json
d = json.load(open("gpurun_out/iter12_b1_$f.json"))
b = d["breakdown"]
print("min_rows=$f b1 value %.1f img/s e2e %.1f clk %s | %s" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], " ".join("%s %.3f" % (k, v["ms_per_batch"]) for k, v in b.items() if isinstance(v, dict) and k.startswith(("dec", "post")))))
PY
done
