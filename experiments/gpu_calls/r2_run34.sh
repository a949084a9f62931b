#!/bin/bash
# round 2, GPU call 34: filtered K-min kernel in dimension 4..15
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py -m gpu -q 2>&1 | tail -5 | cut -c1-250
echo "== timings"
timeout 900 python experiments/knn_lowd_check.py 500000 3 5 8 12 15 2>&1 | tail -6
