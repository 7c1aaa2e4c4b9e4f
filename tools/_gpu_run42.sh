mkdir -p gpurun_out
n=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/s42_bench_${n}gpu.json 2> gpurun_out/s42_bench_${n}gpu.err; tail -c 300 gpurun_out/s42_bench_${n}gpu.json; echo; tail -2 gpurun_out/s42_bench_${n}gpu.err | cut -c1-200
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $n --steps 2 --warmup 1 | tail -c 300
