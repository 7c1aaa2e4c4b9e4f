for d in 0 1.1e-8; do
  echo "== BC_K3_DEBIAS=$d"
  BC_K3_DEBIAS=$d timeout 300 python tools/k3_check.py --models dmv,imdb0,imdb1,imdb2,imdb3,imdb4 --reps 1 --nq 262144 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    pm=r.get('parity_mixed',{}); pd=r.get('parity_dense',{}); o=r.get('bits_k3_vs_fp64_oracle',{})
    print(r['model'], 'golden cases (BITS+DENSE) max %.2e mean %+.2e | golden cases (DENSE) max %.2e mean %+.2e | 2000 range queries (BITS) max %.2e mean %+.2e' % (pm.get('max_rel',0), pm.get('mean_signed_rel',0), pd.get('max_rel',0), pd.get('mean_signed_rel',0), o.get('max_abs_rel',0), o.get('mean_signed_rel',0)))
"
done
