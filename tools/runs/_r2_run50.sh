#!/bin/bash
# round 2, run 50: ncu launch list of the final bench command; ncu --set full of the final K3 (DENSE rows, IMDB-1) and of bc_spec_bits (Census)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_50_launches_bench.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --sustained-seconds 0.05 --dmv-queries 1e7 > gpurun_out/r2_50_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k3_kernelILi2 -s 1 -c 1 -o gpurun_out/r2_50_k3_dense -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_50_ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:bc_spec_bits -s 6 -c 1 -o gpurun_out/r2_50_spec_census -f python bench.py --steps 2 --warmup 3 --cpu-seconds 0.2 --sustained-seconds 0 --dmv-queries 0 --no-secondary --streams 1 > gpurun_out/r2_50_ncu_spec.log 2>&1; echo "ncu spec rc=$?"
ls -la gpurun_out/r2_50_*.ncu-rep gpurun_out/r2_50_launches_bench.csv
