#!/bin/bash
# round 2, run 21 (8 GPUs): the driver's scaling launch at N = 8, both arms
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_21_bench_8gpu.json 2> gpurun_out/r2_21_bench_8gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_21_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_21_ref_8gpu.json 2> gpurun_out/r2_21_ref_8gpu.err; echo "ref rc=$?"; tail -3 gpurun_out/r2_21_ref_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_21_bench_8gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('inproc', d.get('e2e_inproc'))
print('dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'])
r=json.loads(open('gpurun_out/r2_21_ref_8gpu.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'], r['n_gpus'])
PY
