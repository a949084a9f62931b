#!/bin/bash
mkdir -p gpurun_out
echo "== distributed GPU tests (2 ranks, NCCL)"
timeout 700 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_distributed_2gpu.txt 2>&1; echo "rc=$?"
grep -v "^frame\|ProcessGroupNCCL\|^$" gpurun_out/r2_pytest_distributed_2gpu.txt | tail -40 | cut -c1-400
echo "== reference arm --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_ref_n2.json 2> gpurun_out/r2_ref_n2.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_ref_n2.err; cut -c1-1200 gpurun_out/r2_ref_n2.json
