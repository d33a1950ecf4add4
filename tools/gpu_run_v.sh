#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_long_ctx.py tests/test_gpu_model.py tests/test_gpu_spec.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2v_pytest.log
timeout 900 python tools/ab_decode.py llama-3.1-8b 8 2048 --timeline "attn_group=0" "attn_group=1" "attn_group=0" "attn_group=1" > gpurun_out/r2v_ab.txt 2>&1; tail -12 gpurun_out/r2v_ab.txt
timeout 600 python tools/ab_decode.py llama-3.2-1b 16 256 "attn_group=0" "attn_group=1" > gpurun_out/r2v_ab_1b.txt 2>&1; tail -3 gpurun_out/r2v_ab_1b.txt
