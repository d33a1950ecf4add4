#!/bin/bash
# first runs of the persistent step kernel: bounded by timeouts (every device-side wait is bounded as well)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -25 | tee gpurun_out/r2b_pytest.log
