mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_joblight.py -x -q 2>&1 | tail -6) > gpurun_out/s45_pytest.log; tail -3 gpurun_out/s45_pytest.log
timeout 900 python tools/workload_report.py --only config3 --out gpurun_out/s45_config3.json > gpurun_out/s45_config3.log 2>&1; tail -2 gpurun_out/s45_config3.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/s45_config3.json'))
print(json.dumps(r['config3_job_light'], indent=1))
PY
