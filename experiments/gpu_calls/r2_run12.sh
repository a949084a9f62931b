#!/bin/bash
# round 2, GPU call 12: K-min filter (v2 + private token sub-lists); 8-GPU shard launch-shape sweep on one GPU
mkdir -p gpurun_out
echo "== knn tc tests"
timeout 900 python -m pytest tests/test_knn_tc_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "== D=64 1e6 x 1e6"
KT_STATS=1 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -6
echo "== shard sweep"
timeout 1200 python experiments/shard_sweep.py 2>&1 | tee gpurun_out/r2_shard_sweep.jsonl | cut -c1-600
