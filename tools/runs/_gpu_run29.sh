K3B=1 timeout 300 python tools/k3_check.py --models dmv,imdb0,imdb1,imdb3 --reps 3 2> gpurun_out/s40_err.txt | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception:
        print(l.strip()[:300]); continue
    print(r['model'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if 'k3_qps' in k or 'error' in k}, 'parity', r.get('parity_mixed',{}).get('max_rel'), r.get('parity_mixed',{}).get('mean_signed_rel'), 'oracle', r.get('bits_k3_vs_fp64_oracle'))
"
tail -3 gpurun_out/s40_err.txt | cut -c1-300
