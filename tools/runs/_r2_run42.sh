#!/bin/bash
# round 2, run 42: asynchronous PACKED submissions (two batches in flight): test + bench e2e
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -k "packed or sharded or in_process" 2>&1 | tail -6
timeout 400 python bench.py --no-secondary --dmv-queries 1e7 --sustained-seconds 0.2 > gpurun_out/r2_42_bench.json 2> gpurun_out/r2_42_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_42_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_42_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['e2e'])[:1200])
PY
