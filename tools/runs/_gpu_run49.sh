mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool (K3, imdb1 + dmv golden cases, 6k range queries)"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -x -q -k "test_infer_cases_fused_kernel and (imdb1 or dmv)" > gpurun_out/s49_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error|hazard" gpurun_out/s49_$tool.log | head -6
done
