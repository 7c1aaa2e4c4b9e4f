for a in 1 2 1 2; do
  BC_K3_AHEAD=$a timeout 200 python tools/workload_report.py --only config3 2>/dev/null | python -c "
import sys, json
r=json.load(sys.stdin)['config3']
print('ahead=$a', ['%s %.3g' % (b['bn'], b['per_kernel']['fused_tensor_core']['device_factors_per_s']) for b in r['per_bn']])
"
done
