#!/bin/bash
# round 2, run 43: consecutive steps on one / two / three streams
for n in 1 2 3 1 2; do
timeout 300 python bench.py --no-secondary --dmv-queries 0 --sustained-seconds 2 --cpu-seconds 0.5 --streams $n 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $n', 'value', d['value'], 'frac', d['roofline']['frac'], 'sustained', d['sustained']['value'], d['sustained']['sm_mhz_median'], d['sustained']['power_w_max'])"
done
