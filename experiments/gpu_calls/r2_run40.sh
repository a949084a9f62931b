#!/bin/bash
# round 2, GPU calls 40 and 46: full GPU suite (with the TensorDot / Kron / SoftDTW / BSpline / complex cases and the extended reference suite), smoke, bench N=1
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2_pytest_gpu.log | tail -8 | cut -c1-250
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1"
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine']))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
PY
