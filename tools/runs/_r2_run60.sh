#!/bin/bash
# round 2, run 60: compute-sanitizer on the scaled K3 (row-maximum exchange between the epilogue groups)
mkdir -p gpurun_out
OUT=gpurun_out/r2_60_sanitizer_scaled_k3.txt
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "== $tool :: scaled results through K3 (tree shapes, 50 000 queries) + wide trees" >> $OUT
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "scaled_results_through_the_fused or (below_the_fp32_range and shape1)" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned|hazard|Race" | sort | uniq -c | head -8 >> $OUT
done
cat $OUT
