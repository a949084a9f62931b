#!/bin/bash
# round 2, GPU call 1: bench sanity, the reference's own test files through the seam, K-min filter A/B, GPU tests, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,driver_version --format=csv
echo "== bench (reduced size sanity)"
timeout 900 python bench.py --size 60000 --steps 2 --warmup 3 --reps 3 > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_small.err; head -c 3000 gpurun_out/r2_bench_small.json; echo
echo "== reference suite through the seam"
timeout 1800 python tests/reference_suite.py --json gpurun_out/reference_suite.json > gpurun_out/reference_suite.log 2>&1; echo "rc=$?"
tail -n 40 gpurun_out/reference_suite.log | cut -c1-600
echo "== knn filter A/B"
for pipe in 0 1; do
  KT_PIPE=$pipe KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -4
done
KT_PIPE=1 KB200_KNN_TC_NSPLIT=4 KT_TIME=1 timeout 600 python experiments/knn_tc_check.py 1000000 1000000 64 2>&1 | tail -4
echo "== pytest -m gpu"
timeout 3000 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_frontend_gpu.py > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"
tail -n 40 gpurun_out/r2_pytest_gpu.log
echo "== bench N=1"
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 1000 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json
