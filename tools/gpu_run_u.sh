#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2u_pytest.log
