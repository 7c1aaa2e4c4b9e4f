#!/bin/bash
# round 2, run 28: K3 knob sweep (early Lambda loads in; W/B ring depths, tail gap)
run() { echo "== $*"; env "$@" timeout 100 python tools/k3_check.py --models imdb1,imdb3,dmv --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
"; }
timeout 60 python tools/k3_check.py --models imdb1,imdb3 --nq 65536 --skip-bench 2>&1 | cut -c1-200
run A=1
run BC_K3_WSTAGES=6 BC_K3_BSTAGES=4
run BC_K3_WSTAGES=3
run BC_K3_GAP=1200
run BC_K3_GAP=4000
run BC_K3_NO_SKEW=1
