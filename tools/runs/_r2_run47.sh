#!/bin/bash
# round 2, run 47 (8 GPUs): the driver's scaling launch at N = 8 with the final bench, both arms
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_47_bench_8gpu.json 2> gpurun_out/r2_47_bench_8gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_47_bench_8gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_47_ref_8gpu.json 2> gpurun_out/r2_47_ref_8gpu.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_47_bench_8gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'single',d['roofline']['single_stream']['value'])
print('e2e',d['e2e']['value'], d['e2e']['one_synchronous_call_per_step']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('inproc', d.get('e2e_inproc',{}).get('value'), d.get('e2e_inproc',{}).get('one_synchronous_call_per_step'))
print('dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'], 'sustained', d['sustained']['value'])
r=json.loads(open('gpurun_out/r2_47_ref_8gpu.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'])
PY
