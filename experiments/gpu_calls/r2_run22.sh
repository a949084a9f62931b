#!/bin/bash
# round 2, GPU call 22: filtered K-min kernel v2 (min-only pipelined groups)
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py -m gpu -q 2>&1 | tail -4 | cut -c1-250
for cfg in "128 3" "128 2"; do
  set -- $cfg
  echo "== knn_d3 1e6: filtered T=$1 B=$2"
  KB200_KF_T=$1 KB200_KF_B=$2 timeout 600 python experiments/configs_bench.py --only knn_d3 --reps 3 2>&1 | tail -1 | cut -c1-400
done
echo "== shard sweep knn_d3"
ONLY=knn_d3 NS=3,6,9,12 timeout 900 python experiments/shard_sweep.py 2>&1 | cut -c1-500
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^kb200_reduce$" -s 1 -c 1 -f -o gpurun_out/prof_r02_knn_d3_filter \
    python experiments/configs_bench.py --only knn_d3 --once > gpurun_out/ncu_r02_knn_d3_filter.log 2>&1
tail -1 gpurun_out/ncu_r02_knn_d3_filter.log | cut -c1-200
