#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/a_pytest.log
timeout 300 python experiments/shard_sim.py > gpurun_out/a_shard_sim.jsonl 2> gpurun_out/a_shard_sim.err; cat gpurun_out/a_shard_sim.jsonl; tail -3 gpurun_out/a_shard_sim.err
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; cat gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
