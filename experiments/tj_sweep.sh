#!/bin/bash
# streamed-tile length: 256 (default cap) vs 512 rows
for k in gauss_fwd gauss_dx lse_f32 knn_d3; do
for c in "" "TJ=512" "TJ=1024"; do
  echo "== $k [$c]"; KB200_CONFIG="$c" python experiments/configs_bench.py --only $k 2>&1 | tail -1 | cut -c1-20,70-140
done; done
