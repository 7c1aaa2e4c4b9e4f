#!/bin/bash
# round 2, run 08: which instruction between groups of MMAs stalls the issuer (tools/microbench/umma_issue.cu)
mkdir -p gpurun_out
timeout 60 tools/microbench/bin/umma_issue > gpurun_out/r2_08_umma_issue.txt 2>&1; cat gpurun_out/r2_08_umma_issue.txt
