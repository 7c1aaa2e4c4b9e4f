mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8)
timeout 500 python tools/k2_sweep.py --points 100x10,50x10,50x50,100x50 --nq 262144 --out gpurun_out/s55_config4_wide.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['n_cols'], r['card'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k.endswith('_ms') or k.endswith('_error') or k.startswith('fused')})
"
