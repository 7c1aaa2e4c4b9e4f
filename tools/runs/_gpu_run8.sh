mkdir -p gpurun_out
(BC_SPEC_DYNAMIC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "infer_cases or synthetic_batch or sparse_in_lists or size_independent or edge_cases or imdb_query" 2>&1 | tail -5) > gpurun_out/s8_pytest_dynamic.log; tail -2 gpurun_out/s8_pytest_dynamic.log
b() { env "$@" timeout 120 python bench.py --model $M --steps 10 --cpu-seconds 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(r['value']/1e9, 4), 'Gq/s frac', round(r['roofline']['frac'], 4), 'e2e', round(r['e2e']['value']/1e9, 3), 'relerr', r['rel_err_max_vs_fp64_oracle'])
"; }
{
M=census
for v in X=1 BC_SPEC_DYNAMIC=1 "BC_SPEC_DYNAMIC=1 BC_SPEC_SYNC_EVERY=2048" BC_SPEC_SYNC_EVERY=2048 "BC_SPEC_DYNAMIC=1 BC_SPEC_SYNC_EVERY=1024"; do echo "== census $v"; b $v; done
M=dmv
for v in X=1 BC_SPEC_DYNAMIC=1 "BC_SPEC_DYNAMIC=1 BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1"; do echo "== dmv $v"; b $v; done
M=imdb1
for v in X=1 BC_SPEC_DYNAMIC=1; do echo "== imdb1 $v"; b $v; done
} > gpurun_out/s8_spec_dynamic.txt 2>&1
cat gpurun_out/s8_spec_dynamic.txt
echo "== K2 two-CTA"
(BC_K2_UMMA_VARIANT=T timeout 150 python -m pytest tests/test_gpu_parity.py -k "batched_large" -x -q 2>&1 | tail -8) > gpurun_out/s8_pytest_k2_2sm.log; tail -4 gpurun_out/s8_pytest_k2_2sm.log
filt() { python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print({k: (round(v, 9) if isinstance(v, float) else v) for k, v in r.items() if k in ('n_cols', 'card', 'simt_ms', 'umma_ms', 'umma_tflops_alg', 'speedup_umma', 'umma_max_rel_vs_fp64', 'umma_error')})
"; }
for v in T A; do echo "== variant $v"; BC_K2_UMMA_VARIANT=$v timeout 200 python tools/k2_sweep.py --points 10x1000,100x1000 2>&1 | filt; done > gpurun_out/s8_k2_2sm.log 2>&1
echo "== variant T 10x10000" >> gpurun_out/s8_k2_2sm.log
BC_K2_UMMA_VARIANT=T timeout 300 python tools/k2_sweep.py --points 10x10000 2>&1 | filt >> gpurun_out/s8_k2_2sm.log
cat gpurun_out/s8_k2_2sm.log
