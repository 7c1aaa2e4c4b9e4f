#!/bin/bash
# round 2, run 04: K3 clock trace of one tile (BC_K3_TRACE) and B ring depth 8 vs 4
mkdir -p gpurun_out
for bs in 8 4; do
  echo "== BC_K3_BSTAGES=$bs"
  BC_K3_BSTAGES=$bs timeout 300 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:v for k,v in d.items() if k.endswith('k3_qps')})
"
done
BC_K3_TRACE=1 timeout 300 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 0 2>&1 | grep -A12 "K3 trace" | head -60
