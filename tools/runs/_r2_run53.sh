#!/bin/bash
# round 2, run 53: K3 with the next tile's BITS rows / mask words fetched one pass ahead
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or FUSED or k3 or tensor or census_through or tree_shapes" 2>&1 | tail -3
timeout 100 python tools/k3_check.py --models imdb1,imdb3,dmv,imdb0 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"
timeout 100 python tools/k3_check.py --models imdb1 --nq 262144 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  262144:', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
"
