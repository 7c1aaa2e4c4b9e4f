#!/bin/bash
# round 2, run 34: NUMA / affinity probe behind the box-to-box spread of the end-to-end leg
timeout 120 python tools/numa_probe.py 2>&1 | tail -40
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" | head
timeout 200 python bench.py --no-secondary --steps 10 --warmup 3 --cpu-seconds 1 --sustained-seconds 0.1 --dmv-queries 1e7 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e', d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])"
