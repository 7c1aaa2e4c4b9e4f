mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s9_pytest.log; tail -2 gpurun_out/s9_pytest.log
timeout 300 python bench.py > gpurun_out/s9_bench_census.json 2> gpurun_out/s9_bench_census.err
timeout 300 python bench.py --model dmv > gpurun_out/s9_bench_dmv.json 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s9_bench_ref.json 2>/dev/null
python -c "
import json
for f in ['census','dmv']:
    r=json.loads(open('gpurun_out/s9_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(r['value']/1e9,3),'Gq/s frac',round(r['roofline']['frac'],4),'e2e',round(r['e2e']['value']/1e9,3),'p50us',round(r['p50_latency_us_scalar_query'],1), 'cpu', round(r['cpu_baseline']['value']))
"
timeout 300 python tools/large_batch_sweep.py --out gpurun_out/s9_large_batch.jsonl > gpurun_out/s9_large_batch.log 2>&1; tail -2 gpurun_out/s9_large_batch.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s9_launches_census.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/s9_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_spec_bits -s 5 -c 1 -o gpurun_out/s9_spec_bits_census -f python bench.py --steps 3 --cpu-seconds 0.5 > gpurun_out/s9_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_spec_bits -s 5 -c 1 -o gpurun_out/s9_spec_bits_dmv -f python bench.py --model dmv --steps 3 --cpu-seconds 0.5 > gpurun_out/s9_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bc_spec_dense -s 3 -c 1 -o gpurun_out/s9_spec_dense_imdb0 -f python tools/workload_report.py --only config3 --factors 131072 > gpurun_out/s9_ncu_c.log 2>&1
echo "== K2 T KS=4"
BC_K2_UMMA_VARIANT=T BC_K2_UMMA_KS=4 timeout 200 python tools/k2_sweep.py --points 10x1000,100x1000 2>&1 | cut -c1-600 > gpurun_out/s9_k2_T_ks4.log; cat gpurun_out/s9_k2_T_ks4.log | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l[:200]); continue
    print(r['n_cols'], r['card'], r.get('umma_ms'), r.get('umma_tflops_alg'), r.get('umma_max_rel_vs_fp64'))
"
ncu --set full --clock-control none -k regex:k2_umma2_kernel -s 8 -c 1 -o gpurun_out/s9_k2_umma2 -f env BC_K2_UMMA_VARIANT=T python tools/k2_sweep.py --points 10x1000 --reps 1 --oracle-sample 0 > gpurun_out/s9_ncu_d.log 2>&1
tail -2 gpurun_out/s9_ncu_d.log
