#!/bin/bash
# round 2, run 05: K3 per-step clock stamps of one tile
mkdir -p gpurun_out
BC_K3_TRACE=1 timeout 300 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 0 2>&1 | grep -A62 "K3 trace" | head -135
