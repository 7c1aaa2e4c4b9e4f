mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_fit.py -x -q 2>&1 | tail -15) > gpurun_out/s15_pytest_fit.log; tail -3 gpurun_out/s15_pytest_fit.log
timeout 400 python tools/fit_bench.py --out gpurun_out/s15_fit_bench.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['table'][:20], 'kernel_ms', r['kernel_ms'], 'host_ms', r['host_enqueue_ms'], 'GB/s', round(r['roofline']['achieved']), 'frac', round(r['roofline']['frac'],4), 'e2e_s', round(r['e2e_from_pinned_host_s'],4))
"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fit_count --csv --log-file gpurun_out/s15_fit_launches.csv python tools/fit_bench.py --reps 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/s15_fit_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; s=i+1; break
for r in rows[s:]:
    if len(r)>=len(h): print(r[h.index('Grid Size')], r[h.index('Metric Value')], r[h.index('Metric Unit')])
PY
