mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s7_pytest.log; tail -2 gpurun_out/s7_pytest.log
timeout 400 python tools/workload_report.py --out gpurun_out/s7_workload_report.json > gpurun_out/s7_workload_report.log 2>&1
grep -A8 sql_text_batch_api gpurun_out/s7_workload_report.log | head -24
b() { env "$@" timeout 120 python bench.py --model $M --steps 10 --cpu-seconds 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(r['value']/1e9, 4), 'Gq/s frac', round(r['roofline']['frac'], 4), 'e2e', round(r['e2e']['value']/1e9, 3))
"; }
M=census
for v in X=1 BC_SPEC_SYNC_EVERY=512 BC_SPEC_SYNC_EVERY=2048; do echo "== census $v"; b $v; done > gpurun_out/s7_census_sync.txt 2>&1
cat gpurun_out/s7_census_sync.txt
# ncu: imdb1 BITS kernel via bench, DENSE kernel via the workload report
ncu --set full --clock-control none --import-source on -k regex:bc_spec_bits -s 5 -c 1 -o gpurun_out/s7_spec_bits_imdb1 -f python bench.py --model imdb1 --steps 3 --cpu-seconds 0.5 > gpurun_out/s7_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_spec_dense -s 8 -c 1 -o gpurun_out/s7_spec_dense_imdb -f python tools/workload_report.py --factors 65536 > gpurun_out/s7_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_spec_bits -s 5 -c 1 -o gpurun_out/s7_spec_bits_dmv -f python bench.py --model dmv --steps 3 --cpu-seconds 0.5 > gpurun_out/s7_ncu_c.log 2>&1
tail -2 gpurun_out/s7_ncu_b.log
