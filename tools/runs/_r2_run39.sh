#!/bin/bash
# round 2, run 39: TF32 hi by cvt.rna.tf32.f32 (one instruction) against add + mask (two)
for v in cvt base cvt base; do
  cp tools/runs/_variants/lib_$v.so bayescard_b200/libbayescard_b200.so
  echo "== $v"
  timeout 100 python tools/k3_check.py --models imdb1,imdb3 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', d['model'], [d.get(k) for k in ('bits_k3_qps','dense_k3_qps','dense_fan_k3_qps')])
    else: print(l.rstrip()[:200])
"
done
cp tools/runs/_variants/lib_cvt.so bayescard_b200/libbayescard_b200.so
timeout 60 python tools/k3_check.py --models imdb1,dmv --nq 65536 --skip-bench 2>&1 | cut -c1-260
