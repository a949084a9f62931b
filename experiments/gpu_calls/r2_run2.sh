#!/bin/bash
# round 2, GPU call 2: bench sanity + full, reference suite, D=3 configuration sweep
mkdir -p gpurun_out
echo "== bench (reduced size sanity)"
timeout 900 python bench.py --size 60000 --steps 2 --warmup 3 --reps 3 > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_small.err; head -c 6000 gpurun_out/r2_bench_small.json; echo
echo "== reference suite through the seam"
timeout 1800 python tests/reference_suite.py --json gpurun_out/reference_suite.json > gpurun_out/reference_suite.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" gpurun_out/reference_suite.log | tail -5
grep -E "^(FAILED|ERROR)" gpurun_out/reference_suite.log | cut -c1-300 | head -40
echo "== D=3 config sweep"
for c in "" "R=8,B=2" "R=8,B=3" "R=8,T=128,B=4" "R=8,T=128,B=5" "R=6,B=3" "U=4" "U=16"; do
  echo "-- [$c]"; KB200_CONFIG="$c" timeout 300 python experiments/configs_bench.py --only gauss_fwd 2>&1 | tail -1 | cut -c1-400
done
echo "== bench N=1"
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 1000 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json
