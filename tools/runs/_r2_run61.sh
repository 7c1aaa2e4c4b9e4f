#!/bin/bash
# round 2, run 61: last sanity of the in-tree library (smoke + K3 / host-pipeline tests), and racecheck of the UNSCALED K3 on a multi-pass batch
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests -x -q -m gpu -k "fused or packed or scaled or sharded" 2>&1 | tail -2
echo "== racecheck :: unscaled K3, 60 000 queries (several passes per CTA), deep_hub_tail" > gpurun_out/r2_61_racecheck_multipass.txt
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_kernel_tree_shapes and 60000 and deep_hub_tail" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|Race|hazards" | sort | uniq -c | head -8 >> gpurun_out/r2_61_racecheck_multipass.txt
cat gpurun_out/r2_61_racecheck_multipass.txt
