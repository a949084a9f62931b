#!/bin/bash
# K-min D=3 tile configurations at full size (M = N = 1e6)
for c in "" "R=2,G=8,T=256,B=3" "R=2,G=4,T=256,B=3" "R=2,G=8,T=128,B=6" "R=2,G=8,T=256,B=4" "G=8" "G=2" "U=2" "U=8" "TJ=128" "TJ=512"; do
  echo "== [$c]"; KB200_CONFIG="$c" python experiments/configs_bench.py --only knn_d3 2>&1 | tail -1 | cut -c1-20,95-135,280-420
done
