#!/bin/bash
# round 2, run 16 (2 GPUs): multi-GPU test + torchrun bench with the in-process leg
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "in_process_multi_gpu or sharded" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_16_bench_2gpu.json 2> gpurun_out/r2_16_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_16_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_16_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])
print('inproc', d.get('e2e_inproc'))
print('dmv', d['dmv_large_batch']['queries_per_s'], 'secondary', d['secondary']['factors_per_s'])
PY
