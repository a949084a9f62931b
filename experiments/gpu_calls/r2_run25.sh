#!/bin/bash
# round 2, GPU call 25: full GPU suite, bench N=1, ncu of the filtered K-min kernel
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2_pytest_gpu.log | tail -8 | cut -c1-250
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1"
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine']))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
print(json.dumps(d.get('parity',{}).get('knn_d3')))
PY
echo "== ncu knn_d3 filtered"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^kb200_reduce$" -s 1 -c 1 -f -o gpurun_out/prof_r02_knn_d3_filter \
    python experiments/configs_bench.py --only knn_d3 --once > gpurun_out/ncu_r02_knn_d3_filter.log 2>&1
tail -1 gpurun_out/ncu_r02_knn_d3_filter.log | cut -c1-200
