#!/bin/bash
# Round-1 profiling pass (run under gpurun on one B200); outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python experiments/configs_bench.py > gpurun_out/configs_r01.jsonl 2> gpurun_out/configs_r01.err
cat gpurun_out/configs_r01.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
for spec in gauss_fwd:1:1000000 knn_d3:1:1000000 gauss_dx:2:1000000 lse_f32:1:1000000 knn_d64:1:200000 lse_f64:1:200000 ksolve_d16:1:500000; do
  IFS=: read case skip size <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb200_reduce -s $skip -c 1 -f \
      -o gpurun_out/prof_r01_$case python experiments/configs_bench.py --only $case --once --size $size > gpurun_out/ncu_$case.log 2>&1
  tail -2 gpurun_out/ncu_$case.log
done
ls -la gpurun_out
