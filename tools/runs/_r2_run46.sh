#!/bin/bash
# round 2, run 46: the notebook's stability experiment with its own per-n knobs (p, skip_zero_bit), up to 100 columns
mkdir -p gpurun_out
timeout 1200 python tools/stability_experiment.py --out gpurun_out/r2_46_stability.json > gpurun_out/r2_46_stability.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_46_stability.txt
