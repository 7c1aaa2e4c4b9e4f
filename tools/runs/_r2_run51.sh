#!/bin/bash
# round 2, run 51: compute-sanitizer on what changed last: K3 tree shapes, asynchronous PACKED submissions, ShardedModel submissions
mkdir -p gpurun_out
OUT=gpurun_out/r2_51_sanitizer_final.txt
: > $OUT
run() {
  echo "== $1 :: $2" >> $OUT
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned|hazard|Race" | sort | uniq -c | head -12 >> $OUT
}
SHAPES="fused_kernel_tree_shapes and 3000"
ASYNC="packed_submissions_in_flight or sharded_model_same_device or packed_wire_format"
for tool in memcheck racecheck; do
  run $tool "K3 tree shapes (every planner / epilogue branch), 3 000 queries" "$SHAPES"
  run $tool "asynchronous PACKED submissions, ShardedModel, PACKED expansion" "$ASYNC"
done
run synccheck "K3 tree shapes" "$SHAPES"
run memcheck "K3 tree shapes, 60 000 queries (several passes per CTA, tail edges under the next tile)" "fused_kernel_tree_shapes and 60000 and (deep_hub_tail or hub_112 or two_hubs)"
cat $OUT
