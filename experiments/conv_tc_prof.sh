#!/bin/bash
# Timing experiments + ncu capture of the tensor-core Gaussian conv kernel (run under gpurun on one B200).
mkdir -p gpurun_out
python experiments/mufu_peak.py
for mb in 16 32 48 64 96 4000; do
  echo "chunk $mb MB"; KB200_CONV_TC_CHUNK_MB=$mb CT_ONLY=16 python experiments/conv_tc_check.py 1000000 2>&1 | grep CT_ONLY
done
for dbg in 1 2; do
  CT_ONLY=16 CT_DBG=$dbg python experiments/conv_tc_check.py 1000000 2>&1 | grep CT_ONLY
done
CT_ONLY=3 python experiments/conv_tc_check.py 1000000 2>&1 | grep CT_ONLY
CT_ONLY=8 python experiments/conv_tc_check.py 1000000 2>&1 | grep CT_ONLY
python experiments/conv_tc_check.py 20000 1000000 2>&1 | tail -8
