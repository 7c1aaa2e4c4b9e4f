#!/bin/bash
# round 2, run 13: K3 (tail stripped) check + the fp32-range tests
mkdir -p gpurun_out
timeout 120 python tools/k3_check.py --models imdb1,dmv,imdb3 --nq 1048576 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['model'], {k:(v if not isinstance(v,dict) else v.get('max_rel',v.get('max_abs_rel'))) for k,v in d.items() if k.endswith('k3_qps') or k.startswith('parity') or 'oracle' in k or 'error' in k})
    else: print(l.rstrip())
"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "below_the_fp32_range or wsparse_run_of_256" 2>&1 | tail -15
