#!/bin/bash
# round 2, run 31: three producer groups (704 threads, 80 registers) against two (576 threads, 96 registers)
for v in g3 ns500; do
  cp tools/runs/_variants/lib_$v.so bayescard_b200/libbayescard_b200.so
  echo "== $v"
  timeout 60 python tools/k3_check.py --models imdb1,imdb3 --nq 65536 --skip-bench 2>&1 | cut -c1-160
  timeout 100 python tools/k3_check.py --models imdb1,imdb3,dmv,imdb0 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"
done
