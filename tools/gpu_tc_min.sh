#!/bin/bash
# A/B of the verify-batch kernel choices: gpurun --timeout 900 -- 'bash tools/gpu_tc_min.sh'
mkdir -p gpurun_out
for o in "tc_min=16,pv_batch_min=17" "tc_min=6" "tc_min=6,scores_batch_min=2"; do echo "$o"; PS_OPTS=$o timeout 600 python tools/tree_verify_timing.py llama-3.1-8b 4 2048 2 3 4 6 8 12 16 2>&1 | tail -7; done | tee gpurun_out/tc_min.log
