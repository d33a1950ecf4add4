#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_long_ctx.py tests/test_gpu_model.py -x -q 2>&1 | tail -12 | tee gpurun_out/r2k_pytest.log
timeout 600 python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2k_timeline_8b_ctx2048.txt 2>&1; tail -14 gpurun_out/r2k_timeline_8b_ctx2048.txt
