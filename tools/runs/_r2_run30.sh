#!/bin/bash
# round 2, run 30: mbarrier try_wait suspend-time hint A/B (library variants built with -DBC_K3_SUSPEND_NS=...)
for v in ns20000 ns500 ns100; do
  cp tools/runs/_variants/lib_$v.so bayescard_b200/libbayescard_b200.so
  echo "== $v"
  timeout 100 python tools/k3_check.py --models imdb1,imdb3,dmv --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"
done
