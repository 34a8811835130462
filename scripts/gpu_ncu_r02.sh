#!/bin/bash
# ncu evidence of one bench batch (run only after the plain command exited 0; numbers under ncu are never bench values):
#  * launch lists with duration + DRAM bytes for configs[1] (1 box / image) and configs[3] (32 boxes / image)
#  * --set full captures of the top kernels of both
mkdir -p gpurun_out
for b in 1 32; do
  timeout 300 python scripts/ncu_batch.py $b > gpurun_out/ncu_plain_b$b.log 2>&1 || { echo "plain run b$b failed"; tail -5 gpurun_out/ncu_plain_b$b.log; exit 1; }
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/ncu_launches_b$b.csv python scripts/ncu_batch.py $b > gpurun_out/ncu_launches_b$b.log 2>&1
  echo "launch list b$b exit $?"
done
if [ "${FULL:-1}" = "1" ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'gemm2_op16_kernel|encoder_attention_kernel|layernorm_kernel' -c 14 \
    -o gpurun_out/r02_ncu_top -f python scripts/ncu_batch.py 1 > gpurun_out/ncu_top.log 2>&1
  echo "full capture b1 exit $?"
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"${POST_REGEX:-upsample_stats_fast|contour_hull_disk|EpiConvT|tok_gemm|t2i_attention|i2t_attention|keys_ln}" -c 16 \
    -o gpurun_out/r02_ncu_post -f python scripts/ncu_batch.py 32 > gpurun_out/ncu_post.log 2>&1
  echo "full capture b32 exit $?"
fi
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
