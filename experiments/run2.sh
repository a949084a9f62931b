#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in "" "B=4" "T=256,B=2" ; do
  echo "== knn_d3 $c"; KB200_CONFIG="$c" python experiments/configs_bench.py --only knn_d3 2>&1 | tail -1
done
python experiments/configs_bench.py 2>&1 | grep -v knn_d3
