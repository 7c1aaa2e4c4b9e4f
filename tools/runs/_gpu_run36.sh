mkdir -p gpurun_out
timeout 600 python tools/workload_report.py --only config3 --out gpurun_out/s36_config3.json > gpurun_out/s36_config3.log 2>&1; tail -3 gpurun_out/s36_config3.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/s36_config3.json'))['config3']
for b in r['per_bn']:
    print(b['bn'], 'auto %.3g f/s' % b['device_factors_per_s'], {k: '%.3g' % v['device_factors_per_s'] for k,v in b['per_kernel'].items()}, 'e2e %.3g' % b['e2e_host_factors_per_s'], 'err %.2g' % b['max_rel_err_vs_fp64_oracle'], 'TF %.1f' % b['dense_tflops'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 2 -c 1 -o gpurun_out/s36_k3_imdb1_dense_fan -f python tools/workload_report.py --only config3 --factors 262144 > gpurun_out/s36_ncu.log 2>&1; tail -1 gpurun_out/s36_ncu.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k3_kernel|bc_spec" --csv --log-file gpurun_out/s36_launches_config3.csv python tools/workload_report.py --only config3 --factors 262144 > /dev/null 2>&1; tail -3 gpurun_out/s36_launches_config3.csv | cut -c1-300
