#!/bin/bash
# round 2, run 03: first run of the rebuilt K3 (A operand in tensor memory, two producer groups, dedicated epilogue warps,
# TMA-staged DENSE weights): parity + throughput (tools/k3_check.py), guarded by a timeout
mkdir -p gpurun_out
timeout 300 python tools/k3_check.py --models imdb1,dmv,imdb3,census --nq 1048576 > gpurun_out/r2_03_k3_check.txt 2>&1; echo "rc=$?"
tail -c 6000 gpurun_out/r2_03_k3_check.txt
