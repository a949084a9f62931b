#!/bin/bash
# round 2, GPU call 35 (8 GPUs): bench --gpus 8 with the final code (scaling sanity check before the driver's own run)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?"
tail -c 300 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s nsplit %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine'],d['config'].get('nsplit')))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
PY
