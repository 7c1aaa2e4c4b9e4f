cp bayescard_b200/libbayescard_b200.so /tmp/lib_keep.so
for v in one two one two; do
  cp bayescard_b200/lib_alt_$v.bin bayescard_b200/libbayescard_b200.so
  timeout 200 python tools/workload_report.py --only config3 2>/dev/null | python -c "
import sys, json
r=json.load(sys.stdin)['config3']
print('constexpr ahead=$v config3', ['%s %.3g' % (b['bn'], b['per_kernel']['fused_tensor_core']['device_factors_per_s']) for b in r['per_bn']])
"
  timeout 100 python tools/k3_check.py --models imdb0,imdb1 --skip-parity --reps 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('   k3_check 1M', r['model'], round(r['dense_k3_qps'],3), round(r['dense_fan_k3_qps'],3))
"
done
cp /tmp/lib_keep.so bayescard_b200/libbayescard_b200.so
