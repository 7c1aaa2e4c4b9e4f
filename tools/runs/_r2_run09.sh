#!/bin/bash
# round 2, run 09: K3 (simple issuer, no stamps) throughput + one ncu --set full capture of the BITS and DENSE kernels on IMDB-1
mkdir -p gpurun_out
timeout 90 python tools/k3_check.py --models imdb1,dmv,imdb3 --nq 1048576 --skip-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['model'], {k:v for k,v in d.items() if k.endswith('k3_qps')})
    else: print(l.rstrip())
"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 4 -c 1 -o gpurun_out/r2_09_k3_bits -f \
   python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_09_ncu_bits.log 2>&1; echo "ncu bits rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k3_kernel -s 10 -c 1 -o gpurun_out/r2_09_k3_dense_fan -f \
   python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity --reps 1 > gpurun_out/r2_09_ncu_dense.log 2>&1; echo "ncu dense rc=$?"
ls -la gpurun_out/*.ncu-rep
