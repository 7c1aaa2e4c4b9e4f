#!/bin/bash
# round 2, run 49: final verification on one GPU: whole GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_49_ref.json 2>/dev/null; echo "ref rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_49_ref.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'], d.get('config'))"
timeout 600 python bench.py > gpurun_out/r2_49_bench.json 2> gpurun_out/r2_49_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_49_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'single',d['roofline']['single_stream']['value'], 'launches', d['gpu_launches'])
print('e2e',d['e2e']['value'], d['e2e']['one_synchronous_call_per_step']['value'])
print('secondary', d['secondary']['factors_per_s'], d['secondary']['roofline']['frac'], 'dmv', d['dmv_large_batch']['queries_per_s'], 'sustained', d['sustained']['value'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline'].get('rel_err_max_gpu_vs_reference'), 'err', d['rel_err_max_vs_fp64_oracle'], 'config', d['config'])
PY
