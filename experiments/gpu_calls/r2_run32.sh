#!/bin/bash
# round 2, GPU call 32: K-min D=64 shard timings (median of 7), tail split on / off
mkdir -p gpurun_out
timeout 900 python experiments/knn_d64_shards.py 2>&1 | tail -6 | cut -c1-250
echo "== forcing the tail split at every size"
KB200_KNN_TC_TAIL_WAVES=100 timeout 900 python experiments/knn_d64_shards.py 2>&1 | tail -6 | cut -c1-250
