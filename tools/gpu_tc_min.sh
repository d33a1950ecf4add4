#!/bin/bash
mkdir -p gpurun_out
for t in 16 2; do echo "tc_min=$t"; PS_TC_MIN=$t timeout 600 python tools/tree_verify_timing.py llama-3.1-8b 4 2048 2 4 8 12 15 16 2>&1 | tail -6; done | tee gpurun_out/tc_min.log
