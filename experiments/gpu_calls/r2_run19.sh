#!/bin/bash
# round 2, GPU call 19: tensor-core position gradients with shared-memory-aware chunks
mkdir -p gpurun_out
echo "== conv tc gradient tests"
timeout 900 python -m pytest tests/test_conv_tc.py -m gpu -q -k "gradient" 2>&1 | tail -12 | cut -c1-250
echo "== gradient timing D=16 1e6"
timeout 900 python experiments/conv_grad_check.py 1000000 16 2>&1 | tail -6
echo "== gradient timing D=8 1e6"
timeout 900 python experiments/conv_grad_check.py 1000000 8 2>&1 | tail -6
