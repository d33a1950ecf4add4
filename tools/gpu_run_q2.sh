#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2q_pytest.log
timeout 600 python tools/ab_decode.py qwen2-0.5b 24 32 "" "fused=0" "fused=0,ops_graph=0" > gpurun_out/r2q_ab_qwen2.txt 2>&1; tail -5 gpurun_out/r2q_ab_qwen2.txt
