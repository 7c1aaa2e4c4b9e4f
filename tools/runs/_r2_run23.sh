#!/bin/bash
# round 2, run 23: clock trace of one K3 tile (trace build) after the register-accumulator epilogue
BC_K3_TRACE=1 timeout 100 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 0 2>&1 | grep -A50 "K3 trace" | head -104
