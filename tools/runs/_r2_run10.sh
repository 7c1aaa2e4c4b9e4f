#!/bin/bash
# round 2, run 10: K3 with the SIMT tail, leaves in ascending size, 8-deep DENSE weight ring
mkdir -p gpurun_out
timeout 120 python tools/k3_check.py --models imdb1,dmv,imdb3,imdb0,imdb2,imdb4,census --nq 1048576 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['model'], {k:(v if not isinstance(v,dict) else v.get('max_rel',v.get('max_abs_rel'))) for k,v in d.items() if k.endswith('k3_qps') or k.startswith('parity') or 'oracle' in k or 'error' in k})
    else: print(l.rstrip())
"
