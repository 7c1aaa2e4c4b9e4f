#!/bin/bash
# round 2, run 01: GPU tests + both bench arms after the advisor fixes / bench rewrite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_01_smi.txt
lscpu | head -30 > gpurun_out/r2_01_lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_01_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r2_01_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_01_bench_ref.json 2> gpurun_out/r2_01_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_01_bench_ref.json; echo; tail -c 600 gpurun_out/r2_01_bench.err
