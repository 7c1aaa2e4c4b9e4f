mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 300 python tools/k3_check.py --models dmv,imdb0,imdb1,imdb3,imdb4 --reps 3 --skip-parity 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print(r['model'], {k:round(v,3) for k,v in r.items() if 'k3_qps' in k}, r.get('bits_k3_vs_fp64_oracle'))
"
timeout 500 python tools/k2_sweep.py --points 10x100,20x100,30x100,20x50,30x50,10x200 --out gpurun_out/s52_config4_fused.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['n_cols'], r['card'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k.endswith('_ms') or k.endswith('_error') or k=='fused_max_rel_vs_fp64'})
"
