#!/bin/bash
# how much of the configs[1] step does the decoder + post work (auxiliary stream) cost? bench with / without it (no results without)
mkdir -p gpurun_out
for f in 0 1 0 1; do
if [ "$f" = "1" ]; then export YSI_DEV_SKIP_DECODER=1; else unset YSI_DEV_SKIP_DECODER; fi
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/iter13_b1_$f.json 2> gpurun_out/iter13_b1_$f.err; echo "bench b1 skip_decoder=$f exit $?"; tail -c 200 gpurun_out/iter13_b1_$f.err
python - <<PY
import json
d = json.load(open("gpurun_out/iter13_b1_$f.json"))
print("skip_decoder=$f b1 value %.1f img/s e2e %.1f clk %s ms/step %.2f" % (d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["ms_per_step"]))
PY
done
