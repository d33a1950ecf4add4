#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_model.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/p_tl_ctx2048d.txt 2>&1
head -10 gpurun_out/p_tl_ctx2048d.txt; tail -12 gpurun_out/p_tl_ctx2048d.txt
timeout 300 python tools/timeline.py llama-3.1-8b 8 64 > gpurun_out/p_tl_ctx64d.txt 2>&1
head -1 gpurun_out/p_tl_ctx64d.txt; tail -11 gpurun_out/p_tl_ctx64d.txt | head -5
