#!/bin/bash
mkdir -p gpurun_out
echo "== chunked tests"
timeout 900 python -m pytest tests/test_chunked.py -m gpu -q -x 2>&1 | tail -15
echo "== chunked perf"
timeout 900 python experiments/chunked_perf.py 2>&1 | tail -6
echo "== block-sparse: reference binder vs ours"
timeout 900 python experiments/ref_vs_ours_blocksparse.py 2>&1 | grep -v Warning | tail -8
echo "== reference suite (seeded)"
timeout 1800 python tests/reference_suite.py --json gpurun_out/reference_suite.json > gpurun_out/reference_suite.log 2>&1; echo "rc=$?"
grep -E "passed|failed" gpurun_out/reference_suite.log | tail -3 | cut -c1-300
echo "== D=3 sweep"
timeout 1500 python experiments/sweep_d3.py 2>&1 | tail -40
