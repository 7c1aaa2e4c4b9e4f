#!/bin/bash
# round 2, run 40: native job-light path after the planner rewrite (no per-query allocation, one text buffer, parallel merge)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_joblight.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/joblight_native_bench.py --queries 330000 --out gpurun_out/r2_40_joblight_native.json 2>&1 | tail -3
