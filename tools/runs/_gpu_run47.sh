mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s47_pytest.log; tail -3 gpurun_out/s47_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/s47_bench.json 2> gpurun_out/s47_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s47_bench.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g frac %.3f clocks %s launches %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches']))
print('secondary', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['secondary'].items() if k!='workload'})
print('p50 scalar us', d['p50_latency_us_scalar_query'], 'cpu', d['cpu_baseline']['value'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -c 400
