#!/bin/bash
# ncu --set full of the fused 32-block mat-vec launches of one Qwen2-0.5B decode layer (q|k|v, o, gate|up, down) + an lm_head slice
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^ps_k_mv32$|^ps_k_rope_kv$" -s 5 -c 6 -f -o gpurun_out/mv32_kernels python tools/prof_decode.py qwen2-0.5b 4 64 1 > gpurun_out/ncu_mv32.log 2>&1
tail -2 gpurun_out/ncu_mv32.log
