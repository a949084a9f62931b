#!/bin/bash
for a in "200000 200000 64 10" "1000000 1000000 64 10" "100000 100000 100 5 KMin_ArgKMin" "65536 131072 32 3 KMin" "50000 60000 16 10" "40000 50000 128 10"; do
  echo "=== $a"; timeout 300 python experiments/knn_tc_check.py $a 2>&1 | tail -7
done
