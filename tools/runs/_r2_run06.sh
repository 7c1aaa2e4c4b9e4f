#!/bin/bash
# round 2, run 06: MMA + commit round-trip latency
timeout 120 tools/microbench/bin/umma_chain 2>&1 | grep -A8 "# latency"
