#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py tests/test_gpu_long_ctx.py -x -q 2>&1 | tail -3
timeout 600 python tools/step_timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2g_step_tl_8b_ctx2048.txt 2>&1; cat gpurun_out/r2g_step_tl_8b_ctx2048.txt | tail -26
