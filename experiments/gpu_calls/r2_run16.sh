#!/bin/bash
# round 2, GPU call 16: where is the K-min filter's time?  (a) no survivors at all, (b) TMEM loads + release only
mkdir -p gpurun_out
for dbg in 0 1 2; do
  echo "== KT_DBG=$dbg"
  KT_DBG=$dbg KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | grep "KT_TIME" | tail -1
done
echo "== KT_DBG=1 STAGES=3"
KT_STAGES=3 KT_DBG=1 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | grep "KT_TIME" | tail -1
echo "== ncu KT_DBG=2"
KT_DBG=2 timeout 900 ncu --set full --clock-control none -k kb200_knn_filter -c 1 -f -o gpurun_out/prof_r02_knn_filter_dbg2 \
    python experiments/knn_tc_check.py 1000000 1000000 64 > gpurun_out/ncu_r02_knn_filter.log 2>&1
tail -2 gpurun_out/ncu_r02_knn_filter.log
