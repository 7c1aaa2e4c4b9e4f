#!/bin/bash
# round 2, run 02: tcgen05.mma dependent-chain latency / interleaved chains / TMEM load bandwidth (tools/microbench/umma_chain.cu)
mkdir -p gpurun_out
timeout 120 tools/microbench/bin/umma_chain > gpurun_out/r2_02_umma_chain.txt 2>&1; echo "rc=$?"
tail -70 gpurun_out/r2_02_umma_chain.txt
