#!/bin/bash
# round 2, run 45 (2 GPUs): ShardedModel submissions in flight: tests + the in-process leg of the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -k "sharded or in_process or packed" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --dmv-queries 1e7 --sustained-seconds 0.2 > gpurun_out/r2_45_bench_2gpu.json 2> gpurun_out/r2_45_bench_2gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_45_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_45_bench_2gpu.json').read().strip().splitlines()[-1])
print('e2e',d['e2e']['value'], d['e2e']['one_synchronous_call_per_step'])
print('inproc', d.get('e2e_inproc'))
PY
