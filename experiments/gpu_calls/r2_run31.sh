#!/bin/bash
# round 2, GPU call 31: tail split of the K-min tensor-core path; shard sizes of an N-GPU run
mkdir -p gpurun_out
echo "== knn tc tests"
timeout 900 python -m pytest tests/test_knn_tc_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-250
for rows in 1000000 500000 250000 125000; do
  for tail in 1 0; do
    echo "== D=64 rows=$rows tail=$tail"
    KB200_KNN_TC_TAIL=$tail timeout 600 python experiments/knn_tc_check.py $rows 1000000 64 2>&1 | grep "tcgen05 path" | cut -c1-120
  done
done
