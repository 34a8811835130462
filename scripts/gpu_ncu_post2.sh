#!/bin/bash
# full ncu capture of the a6 / a7 kernels and the decoder's per-box kernels of one configs[3] batch (32 boxes / image)
mkdir -p gpurun_out
timeout 300 python scripts/ncu_batch.py 32 > gpurun_out/ncu_plain_b32.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_b32.log; exit 1; }
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"${POST_REGEX:-upsample_stats_fast|contour_hull_disk|keys_ln|EpiGeneric}" -c 12 \
  -o gpurun_out/r02_ncu_post2 -f python scripts/ncu_batch.py 32 > gpurun_out/ncu_post2.log 2>&1
echo "full capture b32 exit $?"
ls -la gpurun_out/*.ncu-rep
