#!/bin/bash
# round 2, run 58: (diagnosis of a hang of the first scaled K3: a per-thread `if (exponent != 0)` guarded warp-wide tcgen05.ld / st
# instructions in the tensor-memory fallback of the epilogue; the probe script ran plain and scaled K3 on single tree shapes with
# 20 s timeouts and is not kept -- tests/test_gpu_parity.py::test_scaled_results_through_the_fused_kernel covers the same shapes)
