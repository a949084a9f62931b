#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in "G=1" "G=2" "G=4" "G=8" "G=4,B=3" "G=4,B=4" "G=8,B=3" "G=4,T=256,B=1" "G=4,T=256,B=2" "R=2,G=8,T=256,B=2" "R=2,G=8,T=128,B=4" "G=4,U=1" ; do
  echo "== $c"; KB200_CONFIG="$c" python experiments/configs_bench.py --only knn_d3 2>&1 | tail -1
done
