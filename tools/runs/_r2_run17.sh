#!/bin/bash
# round 2, run 17: compute-sanitizer sweep (VERDICT r1 missing #7): K2 (single-CTA and two-CTA tcgen05 GEMM), K-spec (self-resetting
# counter ring), K1, bc_convert.cu (RANGE / SPARSE / WSPARSE / PACKED expansion), the rebuilt K3 and the scaled entry points
mkdir -p gpurun_out
OUT=gpurun_out/r2_17_sanitizer.txt
: > $OUT
run() {  # tool, label, pytest -k expression, extra env
  echo "== $1 :: $2" >> $OUT
  env $4 timeout 600 compute-sanitizer --tool $1 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|out of bounds|misaligned|hazard|Race" | sort | uniq -c | head -12 >> $OUT
}
SEL_K3="test_infer_cases_fused_kernel and (imdb1 or dmv)"
SEL_K12="test_infer_cases_both_kernels and (dmv or imdb3)"
SEL_CONV="packed_wire_format or wsparse_rows_expand or wsparse_run_of_256 or sharded_model_same_device"
SEL_K2="test_batched_large_domain_path and (shape0 or shape4)"
SEL_SC="test_results_below_the_fp32_range and (shape1 or shape3)"
for tool in memcheck racecheck; do
  run $tool "K3 (A in tensor memory, 5 roles)" "$SEL_K3" "X=1"
  run $tool "K1 + K-spec" "$SEL_K12" "X=1"
  run $tool "bc_convert.cu + host pipelines" "$SEL_CONV" "X=1"
  run $tool "K2 tcgen05 GEMM, single CTA" "$SEL_K2" "X=1"
  run $tool "K2 tcgen05 GEMM, cta_group::2" "$SEL_K2" "BC_K2_UMMA_VARIANT=T"
  run $tool "scaled results (K1, K2 + row renormalisation)" "$SEL_SC" "X=1"
done
run synccheck "K3" "$SEL_K3" "X=1"
cat $OUT
