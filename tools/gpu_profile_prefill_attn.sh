#!/bin/bash
# ncu evidence for the batched attention kernels of a prefill (2 layers of the 8B shape, 2048 tokens in chunks of 128):
# launch list + `--set full` of the last chunks' scores / P.V kernels, and of the decode scores kernel.
# gpurun --timeout 1500 -- 'bash tools/gpu_profile_prefill_attn.sh'; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 300 python tools/ab_prefill.py 2>&1 | tee gpurun_out/pa_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/pa_launches.csv python tools/prof_prefill.py llama-3.1-8b 2 2048 128 > gpurun_out/pa_ncu.log 2>&1
python tools/launch_list.py gpurun_out/pa_launches.csv > gpurun_out/pa_launches.txt 2>&1; head -12 gpurun_out/pa_launches.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ps_k_attn_pv_tile|ps_k_attn_scores_tile|ps_k_softmax_ext" -s 96 -c 6 -f -o gpurun_out/prefill_attn python tools/prof_prefill.py llama-3.1-8b 2 2048 128 > gpurun_out/pa_ncu_full.log 2>&1
tail -2 gpurun_out/pa_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ps_k_attn1|ps_k_attn2" -s 8 -c 4 -f -o gpurun_out/decode_attn python tools/prof_decode.py llama-3.1-8b 4 2048 1 > gpurun_out/pa_ncu_dec.log 2>&1
tail -2 gpurun_out/pa_ncu_dec.log
ls -la gpurun_out/*.ncu-rep
