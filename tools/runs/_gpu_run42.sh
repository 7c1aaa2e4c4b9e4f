mkdir -p gpurun_out
for n in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/s42_bench_${n}gpu.json 2> gpurun_out/s42_bench_${n}gpu.err; tail -c 300 gpurun_out/s42_bench_${n}gpu.json; echo; tail -2 gpurun_out/s42_bench_${n}gpu.err | cut -c1-200
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/large_batch_sweep.py --sizes 1e9 > gpurun_out/s42_large_batch_8gpu.jsonl 2> gpurun_out/s42_lb.err; tail -2 gpurun_out/s42_large_batch_8gpu.jsonl | cut -c1-600; tail -2 gpurun_out/s42_lb.err | cut -c1-300
