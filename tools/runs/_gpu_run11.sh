mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s11_bench_census_2gpu.json 2> gpurun_out/s11_bench_2gpu.err
tail -c 1500 gpurun_out/s11_bench_census_2gpu.json | python -c "
import sys, json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2gpu', r['n_gpus'], round(r['value']/1e9,3),'Gq/s e2e', round(r['e2e']['value']/1e9,3), r['scaling'])
" || tail -5 gpurun_out/s11_bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/s11_bench_ref_2gpu.json 2>> gpurun_out/s11_bench_2gpu.err; tail -c 300 gpurun_out/s11_bench_ref_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/large_batch_sweep.py --sizes 1e8,1e9 --out gpurun_out/s11_large_batch_2gpu.jsonl > gpurun_out/s11_large_batch_2gpu.log 2>&1; tail -2 gpurun_out/s11_large_batch_2gpu.log | cut -c1-300
(timeout 300 python -m pytest tests/test_gpu_parity.py -k "replicas or sharded" -x -q 2>&1 | tail -3)
