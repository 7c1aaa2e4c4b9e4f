mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_fit.py -x -q 2>&1 | tail -15) > gpurun_out/s14_pytest_fit.log; tail -3 gpurun_out/s14_pytest_fit.log
timeout 400 python tools/fit_bench.py --out gpurun_out/s14_fit_bench.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['table'][:20], 'kernel_ms', r['kernel_ms'], 'GB/s', round(r['roofline']['achieved']), 'frac', round(r['roofline']['frac'],4), 'e2e_s', round(r['e2e_from_pinned_host_s'],4))
"
ncu --set full --clock-control none --import-source on -k regex:fit_count_kernel -s 4 -c 1 -o gpurun_out/s14_fit_count_dmv -f python tools/fit_bench.py --reps 3 > gpurun_out/s14_ncu.log 2>&1; tail -1 gpurun_out/s14_ncu.log
echo "== K2 two-CTA, relaxed arrive"
(BC_K2_UMMA_VARIANT=T timeout 150 python -m pytest tests/test_gpu_parity.py -k "batched_large" -x -q 2>&1 | tail -3)
for ks in 2 4; do BC_K2_UMMA_VARIANT=T BC_K2_UMMA_KS=$ks timeout 200 python tools/k2_sweep.py --points 10x1000,100x1000 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l[:200]); continue
    print('T ks=$ks', r['n_cols'], r['card'], r.get('umma_ms'), r.get('umma_tflops_alg'), r.get('umma_max_rel_vs_fp64'))
"; done 2>&1 | tee gpurun_out/s14_k2_T.txt
