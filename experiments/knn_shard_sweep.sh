#!/bin/bash
# K-min kernel at the row count of one GPU's shard (8-GPU run: 125000 queries against 1e6 points)
for rows in 125000 250000; do
for c in "" "T=128,B=4" "R=2,G=8,T=256,B=2" "R=2,G=8,T=256,B=3" "R=2,G=8,T=128,B=4" "R=1,G=8,T=256,B=4" "R=2,G=4,T=256,B=2"; do
  for ts in 0 1; do
  echo "== rows=$rows cfg=[$c] tailsplit=$ts"; KB200_TAILSPLIT=$ts KB200_CONFIG="$c" python experiments/configs_bench.py --only knn_d3 --rows $rows 2>&1 | tail -1 | cut -c1-60,180-400
  done
done
done
