#!/bin/bash
# ncu of one configs[3] batch (32 boxes / image): launch list (duration + DRAM bytes) and a full capture of the decoder / post kernels
mkdir -p gpurun_out
timeout 300 python scripts/ncu_batch.py 32 > gpurun_out/ncu_plain_b32.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_b32.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/ncu_launches_b32.csv python scripts/ncu_batch.py 32 > gpurun_out/ncu_launches_b32.log 2>&1
echo "launch list b32 exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"${POST_REGEX:-upsample_stats_fast|contour_hull_disk|EpiConvT|t2i_attention_rows|i2t_attention4|keys_ln}" -c 14 \
  -o gpurun_out/r02_ncu_post -f python scripts/ncu_batch.py 32 > gpurun_out/ncu_post.log 2>&1
echo "full capture b32 exit $?"
