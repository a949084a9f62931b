#!/bin/bash
# round 2, GPU call 27: K-min tensor-core filter with 512-row CTAs (KT_RBS=4)
mkdir -p gpurun_out
echo "== knn tc tests"
timeout 900 python -m pytest tests/test_knn_tc_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
for rbs in 4 2; do
  echo "== D=64 1e6 x 1e6 KT_RBS=$rbs"
  KT_RBS=$rbs KT_STATS=1 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -5 | cut -c1-250
done
echo "== D=32 1e6 KT_RBS=4 / 2"
KT_RBS=4 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 32 2>&1 | grep -m1 "KT_TIME\|tcgen05 path" 
KT_RBS=2 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 32 2>&1 | grep -m1 "KT_TIME\|tcgen05 path"
echo "== ncu filter D=64 RBS=4"
KT_RBS=4 timeout 900 ncu --set full --clock-control none --import-source on -k kb200_knn_filter -c 1 -f -o gpurun_out/prof_r02_knn_filter_rbs4 \
    python experiments/knn_tc_check.py 1000000 1000000 64 > gpurun_out/ncu_r02_knn_filter.log 2>&1
tail -2 gpurun_out/ncu_r02_knn_filter.log | cut -c1-200
