mkdir -p gpurun_out
timeout 400 python tools/workload_report.py --out gpurun_out/s5_workload_report.json > gpurun_out/s5_workload_report.log 2>&1
b() { env "$@" timeout 120 python bench.py --model $M --steps 10 --cpu-seconds 1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(r['value']/1e9, 4), 'Gq/s frac', round(r['roofline']['frac'], 4), 'e2e', round(r['e2e']['value']/1e9, 3))
"; }
for M in dmv imdb1; do
  echo "== $M default"; b X=1
  echo "== $M SYNC=1024"; b BC_SPEC_SYNC_EVERY=1024
  echo "== $M SYNC=1024 T256 B1"; b BC_SPEC_SYNC_EVERY=1024 BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1
  echo "== $M SYNC=256 T256 B1"; b BC_SPEC_SYNC_EVERY=256 BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1
  echo "== $M T256 B1"; b BC_SPEC_THREADS=256 BC_SPEC_MIN_BLOCKS=1
done > gpurun_out/s5_spec_sync_variants.txt 2>&1
cat gpurun_out/s5_spec_sync_variants.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s5_k2_launches_100x1000.csv python tools/k2_sweep.py --points 100x1000 --reps 1 --oracle-sample 0 > gpurun_out/s5_k2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_umma_kernel -s 8 -c 1 -o gpurun_out/s5_k2_umma_10x1000 -f python tools/k2_sweep.py --points 10x1000 --reps 1 --oracle-sample 0 > gpurun_out/s5_k2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_gemm_simt -s 4 -c 1 -o gpurun_out/s5_k2_simt_10x1000 -f python tools/k2_sweep.py --points 10x1000 --reps 1 --oracle-sample 0 >> gpurun_out/s5_k2_ncu.log 2>&1
tail -3 gpurun_out/s5_k2_ncu.log
