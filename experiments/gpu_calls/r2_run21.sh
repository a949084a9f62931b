#!/bin/bash
# round 2, GPU call 21: filtered K-min kernel: tests + ncu
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py -m gpu -q 2>&1 | tail -4 | cut -c1-250
echo "== ncu knn_d3 filtered"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^kb200_reduce$" -s 1 -c 1 -f -o gpurun_out/prof_r02_knn_d3_filter \
    python experiments/configs_bench.py --only knn_d3 --once > gpurun_out/ncu_r02_knn_d3_filter.log 2>&1
tail -2 gpurun_out/ncu_r02_knn_d3_filter.log | cut -c1-200
