#!/bin/bash
# round 2, run 44 (2 GPUs): the driver's launch with the final bench (two streams, two batches in flight), both arms
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_44_bench_2gpu.json 2> gpurun_out/r2_44_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_44_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_44_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'single',d['roofline']['single_stream'])
print('e2e',d['e2e']['value'], d['e2e']['one_synchronous_call_per_step'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('inproc', d.get('e2e_inproc',{}).get('value'), 'dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'], 'sustained', d['sustained']['value'])
PY
timeout 400 python bench.py > gpurun_out/r2_44_bench_1gpu.json 2>/dev/null; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_44_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'single',d['roofline']['single_stream'])
print('e2e',d['e2e']['value'], d['e2e']['one_synchronous_call_per_step'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'], 'sustained', d['sustained']['value'], 'clocks', d['clocks'])
PY
