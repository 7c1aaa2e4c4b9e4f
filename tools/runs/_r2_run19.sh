#!/bin/bash
# round 2, run 19: native job-light path on the GPU: parity test + throughput against the Python path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_joblight.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/joblight_native_bench.py --queries 330000 --out gpurun_out/r2_19_joblight_native.json 2>&1 | tail -3
