#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -3
for u in 1 0; do
timeout 600 python tools/timeline.py llama-3.1-8b 8 2048 --opt=rw_unroll2=$u > gpurun_out/r2l_timeline_u$u.txt 2>&1; echo "unroll2=$u"; head -1 gpurun_out/r2l_timeline_u$u.txt; grep -A10 "per-kernel-kind" gpurun_out/r2l_timeline_u$u.txt | head -10
done
