#!/bin/bash
# round 2, GPU call 23: filtered K-min kernel: CTA shape sweep (rows per thread, tile, residency)
mkdir -p gpurun_out
for cfg in "128 3 4 1024" "128 5 2 512" "256 2 2 1024" "128 4 2 512" "256 3 2 512" "128 6 2 512" "128 5 2 1024"; do
  set -- $cfg
  echo "== knn_d3 1e6: filtered T=$1 B=$2 R=$3 TJ=$4"
  KB200_KF_T=$1 KB200_KF_B=$2 KB200_KF_R=$3 KB200_KF_TJ=$4 timeout 600 python experiments/configs_bench.py --only knn_d3 --reps 3 2>&1 | tail -1 | cut -c1-330
done
