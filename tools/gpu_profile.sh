#!/bin/bash
# ncu evidence for profiles/: launch list of a decode run + `--set full` of the six kernels of one 8B layer and an lm_head slice,
# and of the (opt-in) persistent step kernel.  Run on a GPU box (gpurun -- 'bash tools/gpu_profile.sh'); numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_8b.csv python bench.py --prompt 1 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/launch_list.py gpurun_out/launches_8b.csv > gpurun_out/launches_8b.txt 2>&1; head -14 gpurun_out/launches_8b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ps_k_rw_matvec|ps_k_attn1|ps_k_attn2" -s 18 -c 7 -f -o gpurun_out/decode_kernels python tools/prof_decode.py llama-3.1-8b 4 2048 1 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
