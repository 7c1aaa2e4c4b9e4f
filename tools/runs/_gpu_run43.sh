mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s43_pytest.log; tail -3 gpurun_out/s43_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python tools/k2_sweep.py --points 10x10,10x100,20x100,10x200,30x100,20x50 --out gpurun_out/s43_config4_fused.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['n_cols'], r['card'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k.endswith('_ms') or k.endswith('_error') or 'fused' in k})
"
