#!/bin/bash
# Round-2 measurement pass (one B200, under gpurun): GPU tests, bench, ncu launch list of the bench command, one
# `ncu --set full` capture per BASELINE kernel at the benched shape.  Outputs land in gpurun_out/; the summaries that
# are judged are copied to profiles/ afterwards (experiments/ncu_summary.py).
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"
tail -n 4 gpurun_out/r2_pytest_gpu.log | cut -c1-200
echo "== bench N=1"
timeout 1500 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
tail -c 400 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print("value %.4e ms %.2f frac %.3f e2e %.4e engine %s"%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['config']['engine']))
for k in d['legs']:
    v=d[k]; print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms','fwd_ms','grad_x_ms','grad_y_ms','ms_per_iteration','frac_of_roofline','grad_x_frac_of_roofline','engine')})
PY
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu --set full per kernel"
# case : kernel regex : launches to skip : size
for spec in "gauss_fwd:^kb200_reduce$:1:1000000" "gauss_dx:^kb200_reduce$:1:1000000" "gauss_dy:^kb200_reduce$:1:1000000" \
            "lse_f32:kb200_reduce:1:1000000" "lse_f64:kb200_reduce:1:500000" "knn_d3:kb200_reduce:2:1000000" \
            "knn_d64:kb200_knn_select:1:1000000" "ksolve_d16:kb200_conv_tc:1:1000000"; do
  IFS=: read case kern skip size <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s $skip -c 1 -f \
      -o gpurun_out/prof_r02_$case python experiments/configs_bench.py --only $case --once --size $size > gpurun_out/ncu_r02_$case.log 2>&1
  tail -1 gpurun_out/ncu_r02_$case.log | cut -c1-200
done
echo "== ncu conv_tc_hd (Gaussian conv D = 64)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb200_conv_tc_hd -s 1 -c 1 -f -o gpurun_out/prof_r02_conv_tc_hd_d64 \
    env CT_ONLY=64 python experiments/conv_tc_check.py 500000 > gpurun_out/ncu_r02_conv_tc_hd.log 2>&1
tail -2 gpurun_out/ncu_r02_conv_tc_hd.log | cut -c1-200
ls -la gpurun_out/prof_r02_* | awk '{print $5, $9}'
