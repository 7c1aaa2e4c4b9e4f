#!/bin/bash
# round 2, run 07: K3 with the software-pipelined issuer + 32-column epilogue loads
mkdir -p gpurun_out
timeout 60 python tools/k3_check.py --models imdb1,dmv,imdb3 --nq 1048576 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['model'], {k:v for k,v in d.items() if k.endswith('k3_qps') or k.startswith('parity') or 'oracle' in k})
    else: print(l.rstrip())
"
BC_K3_TRACE=1 timeout 60 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 0 2>&1 | grep -A9 "K3 trace" | head -22
