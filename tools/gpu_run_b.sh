#!/bin/bash
mkdir -p gpurun_out
# ctx 2048 prefill uses table ops (many launches): skip them by kernel-name filter
ncu --set full --clock-control none --import-source on -k regex:ps_k_matvec_q4k_tma -s 4 -c 4 -f -o gpurun_out/b_matvec python tools/prof_decode.py llama-3.1-8b 4 64 3 > gpurun_out/b_ncu_matvec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ps_k_attn[12] -s 8 -c 2 -f -o gpurun_out/b_attn python tools/prof_decode.py llama-3.1-8b 4 2048 3 > gpurun_out/b_ncu_attn.log 2>&1
tail -3 gpurun_out/b_ncu_matvec.log gpurun_out/b_ncu_attn.log
