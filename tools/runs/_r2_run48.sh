#!/bin/bash
# round 2, run 48: three producer groups again, with the skipped-block barrier waits that keep every group in phase
for v in g3 g2; do
  cp tools/runs/_variants/lib_$v.so bayescard_b200/libbayescard_b200.so
  echo "== $v"
  timeout 40 python tools/k3_check.py --models imdb1,imdb3 --nq 65536 --skip-bench 2>&1 | cut -c1-140
  timeout 60 python tools/k3_check.py --models imdb1,imdb3,dmv --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"
done
