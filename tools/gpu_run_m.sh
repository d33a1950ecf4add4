#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/timeline.py llama-3.1-8b 8 64 --cta > gpurun_out/m_tl_v2.txt 2>&1
head -16 gpurun_out/m_tl_v2.txt | tail -14; tail -32 gpurun_out/m_tl_v2.txt
