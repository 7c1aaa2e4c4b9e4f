#!/bin/bash
# round 2, run 14: fp32-range tests (K2 leaf fix), PACKED tests, bench with the PACKED e2e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "below_the_fp32_range or packed or batched_large_domain" 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --dmv-queries 1e8 > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_14_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_14_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',json.dumps(d['e2e'])[:900])
print('secondary', d['secondary']['factors_per_s'], d['secondary']['roofline']['frac'], d['secondary']['e2e_factors_per_s_wsparse_host'])
PY
for c in 65536 131072 524288 1048576; do echo "chunk $c"; BC_PACKED_CHUNK=$c timeout 200 python bench.py --steps 5 --warmup 3 --dmv-queries 0 --no-secondary --cpu-seconds 0.5 --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['h2d_gbs_per_rank'], d['e2e']['h2d_peak_gbs_per_rank'])"; done
