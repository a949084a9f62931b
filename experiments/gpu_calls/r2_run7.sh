#!/bin/bash
# round 2, GPU call 7: reworked K-min filter (two threads per row, shared row state, tokens, streamed k-blocks for large D)
mkdir -p gpurun_out
echo "== knn tc tests"
timeout 900 python -m pytest tests/test_knn_tc_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "== D=64 1e6 x 1e6"
KT_STATS=1 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6
echo "== D=64 1e4 x 1e6 (column splits)"
KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 10000 1000000 64 2>&1 | tail -6
echo "== D=784 70000 x 70000"
KT_STATS=1 KT_TIME=1 timeout 900 python experiments/knn_tc_check.py 70000 70000 784 2>&1 | tail -8
echo "== D=256 200000 x 200000"
KT_TIME=1 KT_FULLCHECK=1 timeout 900 python experiments/knn_tc_check.py 200000 200000 256 2>&1 | tail -8
echo "== ncu filter D=64"
timeout 900 ncu --set full --clock-control none --import-source on -k kb200_knn_filter -c 1 -f -o gpurun_out/prof_r02_knn_filter \
    python experiments/knn_tc_check.py 1000000 1000000 64 > gpurun_out/ncu_r02_knn_filter.log 2>&1
tail -3 gpurun_out/ncu_r02_knn_filter.log
