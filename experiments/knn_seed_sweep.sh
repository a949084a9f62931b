#!/bin/bash
for rows in 125000 250000 500000 1000000; do
  for pen in 0.02 0.008 0.004 0.001; do
  echo "== rows=$rows penalty=$pen"; KB200_SPLIT_PENALTY=$pen python experiments/configs_bench.py --only knn_d3 --rows $rows 2>&1 | tail -1 | cut -c1-60,90-140,300-420
  done
done
