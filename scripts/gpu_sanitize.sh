#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): usage  scripts/gpu_sanitize.sh racecheck|synccheck|memcheck
TOOL=${1:-racecheck}
mkdir -p gpurun_out
for c in attn tower tower80; do
  timeout 900 python scripts/sanitize_cases.py $c > gpurun_out/san_plain_$c.log 2>&1 || { echo "plain $c failed"; tail -5 gpurun_out/san_plain_$c.log; continue; }
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 40 python scripts/sanitize_cases.py $c > gpurun_out/san_${TOOL}_$c.log 2>&1
  echo "$TOOL $c exit $?"; grep -E "SUMMARY|done|Race|hazard|Error" gpurun_out/san_${TOOL}_$c.log | tail -8
done
