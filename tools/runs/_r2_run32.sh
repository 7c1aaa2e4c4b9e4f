#!/bin/bash
# round 2, run 32: three producer groups: where does it stop?
cp tools/runs/_variants/lib_g3.so bayescard_b200/libbayescard_b200.so
for nq in 131072 262144 1048576; do
for cfg in "A=1" "BC_K3_NO_SKEW=1"; do
  echo "== g3 nq=$nq $cfg"
  env $cfg timeout 40 python tools/k3_check.py --models imdb1 --nq $nq --skip-parity --reps 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"; echo "rc=$?"
done; done
