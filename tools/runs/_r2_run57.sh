#!/bin/bash
# round 2, run 57: scaled results through K3 (exponent carried in the epilogue) -- short timeouts
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -k "below_the_fp32_range or scaled_results_through_the_fused" 2>&1 | tail -12
