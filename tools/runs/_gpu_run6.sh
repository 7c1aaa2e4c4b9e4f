mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -k "batched_large" -x -q 2>&1 | tail -5) > gpurun_out/s6_pytest_k2.log; tail -3 gpurun_out/s6_pytest_k2.log
filt() { python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: (round(v, 9) if isinstance(v, float) else v) for k, v in r.items() if k in ('n_cols', 'card', 'simt_ms', 'umma_ms', 'umma_tflops_alg', 'speedup_umma', 'umma_max_rel_vs_fp64', 'umma_error')})
"; }
for v in A B; do echo "== variant $v"; BC_K2_UMMA_VARIANT=$v timeout 200 python tools/k2_sweep.py --points 10x1000,100x1000 2>&1 | filt; done > gpurun_out/s6_k2_variants.log 2>&1
for v in C D; do echo "== variant $v"; BC_K2_UMMA_VARIANT=$v timeout 200 python tools/k2_sweep.py --points 10x100,100x100 2>&1 | filt; done >> gpurun_out/s6_k2_variants.log 2>&1
echo "== default 10x10000" >> gpurun_out/s6_k2_variants.log
timeout 300 python tools/k2_sweep.py --points 10x10000 2>&1 | filt >> gpurun_out/s6_k2_variants.log
cat gpurun_out/s6_k2_variants.log
