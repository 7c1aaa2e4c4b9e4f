#!/bin/bash
# round 2, run 29: ncu --set full + source-level stall samples of the current K3 (IMDB-1, DENSE rows)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 3 -c 1 -o gpurun_out/r2_29_k3 -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_29_ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
ls -la gpurun_out/r2_29_k3.ncu-rep
