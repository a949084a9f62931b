#!/bin/bash
# round 2, GPU call 24: filtered K-min kernel v3 (unrolled rescan): tests + shapes + 8-GPU shard
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py -m gpu -q 2>&1 | tail -3 | cut -c1-250
for cfg in "128 3 4 1024" "128 5 2 512"; do
  set -- $cfg
  echo "== knn_d3 1e6: filtered T=$1 B=$2 R=$3 TJ=$4"
  KB200_KF_T=$1 KB200_KF_B=$2 KB200_KF_R=$3 KB200_KF_TJ=$4 timeout 600 python experiments/configs_bench.py --only knn_d3 --reps 3 2>&1 | tail -1 | cut -c1-330
  KB200_KF_T=$1 KB200_KF_B=$2 KB200_KF_R=$3 KB200_KF_TJ=$4 ONLY=knn_d3 SWEEP=0 WORLDS=1,2,4,8 timeout 600 python experiments/shard_sweep.py 2>&1 | cut -c1-300
done
