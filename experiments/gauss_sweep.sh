#!/bin/bash
# Gaussian conv forward (D=3) tile configurations at full size
for c in "" "U=16" "U=4" "T=128,B=8" "T=512,B=2" "R=8,B=2" "R=8,B=3" "R=6,B=3" "R=2,B=6" "TJ=256" "B=5" "B=3"; do
  echo "== [$c]"; KB200_CONFIG="$c" python experiments/configs_bench.py --only gauss_fwd 2>&1 | tail -1 | cut -c1-22,70-140,300-420
done
