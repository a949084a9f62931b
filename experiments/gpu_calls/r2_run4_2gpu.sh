#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
echo "== distributed GPU tests (2 ranks, NCCL)"
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2_pytest_distributed_2gpu.txt
echo "== bench --gpus 2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_n2.err; cat gpurun_out/r2_bench_n2.json | cut -c1-6000
echo "== reference arm --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-1500
