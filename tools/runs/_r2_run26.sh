#!/bin/bash
# round 2, run 26: K3 with the top-of-tree chain folded under the next tile (skewed step sequence): parity first, short timeouts
timeout 60 python tools/k3_check.py --models imdb1 --nq 4096 --skip-bench 2>&1 | tail -2 | cut -c1-400 || echo "TIMEOUT small"
timeout 90 python tools/k3_check.py --models imdb1 --nq 262144 2>&1 | tail -2 | cut -c1-1500 || echo "TIMEOUT imdb1"
timeout 200 python tools/k3_check.py --models dmv,imdb0,imdb2,imdb3,imdb4 --nq 1048576 2>&1 | tail -6 | cut -c1-1500
timeout 100 python tools/k3_check.py --models imdb1 --nq 1048576 --skip-parity 2>&1 | tail -2 | cut -c1-1500
BC_K3_NO_SKEW=1 timeout 100 python tools/k3_check.py --models imdb1,imdb3 --nq 1048576 --skip-parity 2>&1 | tail -3 | cut -c1-1500
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or FUSED or k3 or tensor or census_through" 2>&1 | tail -5
