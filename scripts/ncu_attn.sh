#!/bin/bash
mkdir -p gpurun_out
CMD="python -m pytest tests/test_gpu_attention.py -q -m gpu -x -p no:cacheprovider -k 'True-2-3 or False-50-2'"
eval $CMD > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:encoder_attention -c 2 -o gpurun_out/prof_attn -f \
  python -m pytest tests/test_gpu_attention.py -q -m gpu -x -p no:cacheprovider -k 'True-2-3 or False-50-2' > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn.log
