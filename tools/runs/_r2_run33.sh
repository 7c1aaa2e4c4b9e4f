#!/bin/bash
# round 2, run 33: the final K3 of the round: whole GPU suite, sanitizers on K3, bench, ncu --set full of the DENSE kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
OUT=gpurun_out/r2_33_sanitizer_k3.txt
: > $OUT
SEL_K3="test_infer_cases_fused_kernel and (imdb1 or dmv or imdb3)"
for tool in memcheck racecheck synccheck; do
  echo "== $tool :: K3 (registers epilogue, skewed step sequence)" >> $OUT
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL_K3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned|hazard|Race" | sort | uniq -c | head -12 >> $OUT
done
cat $OUT
timeout 600 python bench.py > gpurun_out/r2_33_bench.json 2> gpurun_out/r2_33_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_33_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_33_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
s=d['secondary']; print('secondary', s['factors_per_s'], s.get('roofline',{}).get('frac'), s.get('max_rel_err_vs_fp64_oracle'))
print('dmv', d['dmv_large_batch']['queries_per_s'])
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 10 -c 1 -o gpurun_out/r2_33_k3_dense -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_33_ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
