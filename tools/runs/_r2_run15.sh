#!/bin/bash
# round 2, run 15: full GPU suite + bench (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_15_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_15_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_15_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'], d['secondary']['roofline']['frac'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('rel_err_max_gpu_vs_reference'))
PY
