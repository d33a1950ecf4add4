#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_long_ctx.py -x -q 2>&1 | tail -6 | tee gpurun_out/r2x_pytest.log
