#!/bin/bash
# round 2, run 36: ncu --set full of the final K3 on DENSE rows (mangled-name filter picks the DENSE_F32 instantiation), bench line
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k3_kernelILi3 -s 1 -c 1 -o gpurun_out/r2_36_k3_dense -f python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_36_ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
ls -la gpurun_out/r2_36_k3_dense.ncu-rep
timeout 600 python bench.py > gpurun_out/r2_36_bench.json 2> gpurun_out/r2_36_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_36_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['h2d_gbs_per_rank'],'frac',d['roofline']['frac'])
s=d['secondary']; print('secondary', s['factors_per_s'], s['roofline']['frac'], s['e2e_factors_per_s_wsparse_host'])
print('dmv', d['dmv_large_batch']['queries_per_s'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_36_ref.json 2>/dev/null; tail -c 600 gpurun_out/r2_36_ref.json
