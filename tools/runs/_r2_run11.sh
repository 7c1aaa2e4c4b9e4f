#!/bin/bash
# round 2, run 11: K3 A/B: SIMT tail on/off x sibling order new/old
for cfg in "" "BC_K3_NO_TAIL=1" "BC_K3_OLD_ORDER=1" "BC_K3_NO_TAIL=1 BC_K3_OLD_ORDER=1"; do
  echo "== cfg: $cfg"
  env $cfg timeout 100 python tools/k3_check.py --models imdb1,dmv,imdb3 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' ', d['model'], {k:v for k,v in d.items() if k.endswith('k3_qps')})
    else: print(l.rstrip())
"
done
