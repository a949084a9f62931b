#!/bin/bash
# round 2, GPU call 20: filtered K-min kernel for D <= 3
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-250
echo "== knn_d3 1e6: generic"
KB200_KMIN_FILTER=0 timeout 600 python experiments/configs_bench.py --only knn_d3 --reps 3 2>&1 | tail -1 | cut -c1-400
for cfg in "128 3" "256 2" "256 1" "128 2"; do
  set -- $cfg
  echo "== knn_d3 1e6: filtered T=$1 B=$2"
  KB200_KF_T=$1 KB200_KF_B=$2 timeout 600 python experiments/configs_bench.py --only knn_d3 --reps 3 2>&1 | tail -1 | cut -c1-400
done
echo "== shard sweep knn_d3"
ONLY=knn_d3 NS=3,6,9,12,19,24 timeout 900 python experiments/shard_sweep.py 2>&1 | cut -c1-500
