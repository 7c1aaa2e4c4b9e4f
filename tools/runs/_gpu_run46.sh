mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s46_pytest.log; tail -3 gpurun_out/s46_pytest.log
timeout 900 python tools/workload_report.py --only config3 --out gpurun_out/s46_config3.json > gpurun_out/s46_config3.log 2>&1; tail -2 gpurun_out/s46_config3.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/s46_config3.json'))
j=r['config3_job_light']; print({k:j[k] for k in ('q_error_50_90_95_99_100','scalar_call_latency_ms_p50_p99','batch_queries_per_s','cpu_port_queries_per_s_1core','max_rel_diff_gpu_vs_cpu_port')})
PY
