#!/bin/bash
# smoke + bench + ncu launch list (run after the parity tests)
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps ${STEPS:-16} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 4000 gpurun_out/bench.json
tail -n 5 gpurun_out/bench.err
