mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_kernel_properties" 2>&1 | tail -15) > gpurun_out/s44_pytest.log; tail -12 gpurun_out/s44_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s44_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/s44_b.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/s44_launches_bench.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; s=i+1; break
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[s:]:
    if len(r)>=len(h):
        k=r[h.index('Kernel Name')][:60]; agg[k][0]+=1; agg[k][1]+=float(r[h.index('Metric Value')])
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print(n, round(t/1e3,1),'us', k)
PY
