#!/bin/bash
# round 2, run 18: stability experiment (notebook generator), missing config-4 grid points, ncu launch list of the bench command,
# ncu --set full of the final K3 (IMDB-1 DENSE + fan-out) and of the PACKED expansion kernel
mkdir -p gpurun_out
timeout 900 python tools/stability_experiment.py --out gpurun_out/r2_18_stability.json > gpurun_out/r2_18_stability.txt 2>&1; echo "stability rc=$?"; cat gpurun_out/r2_18_stability.txt | tail -12
timeout 900 python tools/k2_sweep.py --points 20x10,50x10,20x10000 --out gpurun_out/r2_18_config4_missing_points.jsonl > gpurun_out/r2_18_k2_sweep.txt 2>&1; echo "sweep rc=$?"; tail -4 gpurun_out/r2_18_k2_sweep.txt | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_18_launches_bench.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --sustained-seconds 0.05 --dmv-queries 1e7 > gpurun_out/r2_18_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 10 -c 1 -o gpurun_out/r2_18_k3_dense_fan -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_18_ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 3 -c 1 -o gpurun_out/r2_18_k3_bits -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_18_ncu_k3b.log 2>&1; echo "ncu k3 bits rc=$?"
ls -la gpurun_out/*.ncu-rep
