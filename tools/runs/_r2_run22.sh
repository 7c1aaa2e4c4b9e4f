#!/bin/bash
# round 2, run 22: K3 with the parent's message in the epilogue warps' registers (two epilogue groups): parity + throughput
mkdir -p gpurun_out
timeout 120 python tools/k3_check.py --models imdb1 --nq 262144 2>&1 | tail -3
timeout 200 python tools/k3_check.py --models dmv,imdb0,imdb2,imdb3,imdb4 --nq 1048576 2>&1 | tail -6
timeout 100 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or FUSED or k3 or tensor or census_through" 2>&1 | tail -5
