#!/bin/bash
# round 2, GPU call 26: filtered Min / ArgMin; K-means assignment timing
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_kmin_filter_gpu.py -m gpu -q 2>&1 | tail -4 | cut -c1-250
echo "== K-means assignment"
timeout 600 python experiments/kmeans_check.py 10000000 1000 3 2>&1 | tail -1
timeout 600 python experiments/kmeans_check.py 1000000 1000000 3 2>&1 | tail -1
timeout 600 python experiments/kmeans_check.py 10000000 100 2 2>&1 | tail -1
