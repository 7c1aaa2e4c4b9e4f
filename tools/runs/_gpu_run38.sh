mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/s38_pytest.log; tail -6 gpurun_out/s38_pytest.log
timeout 600 python tools/workload_report.py --only config3 --out gpurun_out/s38_config3.json > gpurun_out/s38_config3.log 2>&1; tail -3 gpurun_out/s38_config3.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/s38_config3.json'))['config3']
for b in r['per_bn']:
    print(b['bn'], 'auto %.3g f/s' % b['device_factors_per_s'], 'e2e wsparse %.3g (%.0f B/factor)' % (b['e2e_host_factors_per_s'], b['e2e_bytes_per_factor']), 'e2e dense %.3g' % b['e2e_host_dense_rows_factors_per_s'], 'err %.2g' % b['max_rel_err_vs_fp64_oracle'])
PY
