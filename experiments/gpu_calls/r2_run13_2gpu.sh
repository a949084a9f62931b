#!/bin/bash
# round 2, GPU call 13 (2 GPUs): distributed GPU test with the reporting harness; bench --gpus 2 with the per-SM launch planner
mkdir -p gpurun_out
echo "== distributed GPU tests (2 ranks, NCCL)"
timeout 500 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_distributed_2gpu.txt 2>&1; echo "rc=$?"
grep -v "^frame\|ProcessGroupNCCL\|TCPStore\|sendBytes\|Exception raised\|^$" gpurun_out/r2_pytest_distributed_2gpu.txt | tail -60 | cut -c1-1500
echo "== bench --gpus 2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"
tail -c 800 gpurun_out/r2_bench_n2.err; cut -c1-3000 gpurun_out/r2_bench_n2.json
