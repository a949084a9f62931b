#!/bin/bash
# round 2, GPU call 18: tensor-core position gradients; the whole GPU suite without -x
mkdir -p gpurun_out
echo "== conv tc tests"
timeout 900 python -m pytest tests/test_conv_tc.py -m gpu -q 2>&1 | tail -12 | cut -c1-250
echo "== gradient timing D=16 1e6"
timeout 900 python experiments/conv_grad_check.py 1000000 16 2>&1 | tail -6
echo "== gradient timing D=32 5e5"
timeout 900 python experiments/conv_grad_check.py 500000 32 2>&1 | tail -6
echo "== pytest -m gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2_pytest_gpu.log | tail -15 | cut -c1-250
