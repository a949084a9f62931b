#!/bin/bash
# round 2, GPU call 14 (2 GPUs): distributed GPU test after the context fix
mkdir -p gpurun_out
echo "== distributed GPU tests (2 ranks, NCCL)"
timeout 500 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_distributed_2gpu.txt 2>&1; echo "rc=$?"
grep -v "^frame\|ProcessGroupNCCL\|TCPStore\|sendBytes\|Exception raised\|^$" gpurun_out/r2_pytest_distributed_2gpu.txt | tail -30 | cut -c1-2500
