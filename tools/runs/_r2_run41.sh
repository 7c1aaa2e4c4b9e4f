#!/bin/bash
# round 2, run 41: K3 on synthetic tree shapes that walk every planner / epilogue branch
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --tb=short -k "fused_kernel_tree_shapes" 2>&1 | tail -40
