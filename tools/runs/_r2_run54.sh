#!/bin/bash
# round 2, run 54: A/B on one box: next-tile prefetch on / off
for rep in 1 2; do
for cfg in "A=1" "BC_K3_NO_PREFETCH=1"; do
  echo "== $cfg"
  env $cfg timeout 100 python tools/k3_check.py --models imdb1,imdb3 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
"
  env $cfg timeout 100 python tools/k3_check.py --models imdb1 --nq 262144 --skip-parity --reps 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   262144', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
"
done; done
