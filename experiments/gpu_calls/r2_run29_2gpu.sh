#!/bin/bash
# round 2, GPU call 29 (2 GPUs): distributed GPU test + bench --gpus 2 with the final code
mkdir -p gpurun_out
echo "== distributed GPU tests (2 ranks, NCCL)"
timeout 500 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_distributed_2gpu.txt 2>&1; echo "rc=$?"
grep -E "passed|failed" gpurun_out/r2_pytest_distributed_2gpu.txt | tail -2
echo "== bench --gpus 2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine']))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
print("cpu_baseline" in d, d.get("cpu_baseline",{}).get("cores"))
PY
