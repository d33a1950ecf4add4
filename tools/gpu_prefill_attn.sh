#!/bin/bash
# Prefill attention A/B + launch list: gpurun --timeout 1500 -- 'bash tools/gpu_prefill_attn.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pa_pytest.log
timeout 300 python tools/ab_prefill.py 2>&1 | tee gpurun_out/pa_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/pa_launches.csv python tools/prof_prefill.py llama-3.1-8b 2 2048 128 > gpurun_out/pa_ncu.log 2>&1
python tools/launch_list.py gpurun_out/pa_launches.csv > gpurun_out/pa_launches.txt 2>&1; head -16 gpurun_out/pa_launches.txt
