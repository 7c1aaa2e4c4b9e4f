#!/bin/bash
# round 2, run 35: host pipelines without the serialising stream wait (H2D of chunk i+1 used to wait for D2H of chunk i): chunk sweep
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "packed or sparse or wsparse or host" 2>&1 | tail -3
for c in 1048576 524288 262144 131072 65536; do echo "chunk $c"; BC_PACKED_CHUNK=$c timeout 200 python bench.py --steps 10 --warmup 3 --dmv-queries 0 --no-secondary --cpu-seconds 0.5 --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'], 'csr', d['e2e']['sparse_csr']['value'])"; done
