#!/bin/bash
mkdir -p gpurun_out
for p in 0 2; do
CT_POLY=$p CT_ONLY=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb200_conv_tc -s 1 -c 1 -f \
   -o gpurun_out/prof_conv_tc_d8_poly$p python experiments/conv_tc_check.py 262144 > gpurun_out/ncu_conv_tc_p$p.log 2>&1
tail -2 gpurun_out/ncu_conv_tc_p$p.log
done
