#!/bin/bash
mkdir -p gpurun_out
echo "== prescale + chunked tests"
timeout 1200 python -m pytest tests/test_prescale_gpu.py tests/test_chunked.py -m gpu -q 2>&1 | tail -25
echo "== full-size tests"
timeout 1800 python -m pytest tests/test_full_size_gpu.py -m gpu -q 2>&1 | tail -25
echo "== bench N=1"
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine']))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
print(json.dumps(d.get('parity'))[:1500])
PY
