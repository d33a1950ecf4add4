#!/bin/bash
# attention kernels with pre-wait loads: parity, then timeline at ctx 2048 and ctx 64
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_long_ctx.py tests/test_gpu_model.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2n_pytest.log
timeout 600 python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2n_timeline_8b_ctx2048.txt 2>&1; head -16 gpurun_out/r2n_timeline_8b_ctx2048.txt; grep -A10 "per-kernel-kind" gpurun_out/r2n_timeline_8b_ctx2048.txt | head -11
timeout 600 python tools/timeline.py llama-3.1-8b 8 64 > gpurun_out/r2n_timeline_8b_ctx64.txt 2>&1; grep -A10 "per-kernel-kind" gpurun_out/r2n_timeline_8b_ctx64.txt | head -11
