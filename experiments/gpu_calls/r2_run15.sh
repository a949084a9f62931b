#!/bin/bash
# round 2, GPU call 15: K-min filter with helper warps (mailboxes)
mkdir -p gpurun_out
echo "== knn tc tests"
timeout 900 python -m pytest tests/test_knn_tc_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "== D=64 1e6 x 1e6"
KT_STATS=1 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6
echo "== D=64 1e6 x 1e6, T0=16"
KT_T0=16 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6 | head -3
echo "== D=64 1e6 x 1e6, T0=256"
KT_T0=256 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6 | head -3
echo "== D=64 1e6 x 1e6, no helper"
KT_HELPER=0 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6 | head -3
echo "== D=64 1e4 x 1e6 (column splits)"
KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 10000 1000000 64 2>&1 | tail -6
echo "== D=100 300000^2"
KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 300000 300000 100 2>&1 | tail -4
echo "== ncu filter D=64"
timeout 900 ncu --set full --clock-control none --import-source on -k kb200_knn_filter -c 1 -f -o gpurun_out/prof_r02_knn_filter_v6 \
    python experiments/knn_tc_check.py 1000000 1000000 64 > gpurun_out/ncu_r02_knn_filter.log 2>&1
tail -3 gpurun_out/ncu_r02_knn_filter.log
