#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dbg_attn_group.py llama-3.1-8b 1 2100 2100 2049 100 2100 2100 2>&1 | tail -12 | tee gpurun_out/r2w_dbg.txt
timeout 600 python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2w_timeline.txt 2>&1; head -12 gpurun_out/r2w_timeline.txt; grep -A10 "per-kernel-kind" gpurun_out/r2w_timeline.txt | head -11
