#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sessions.py tests/test_gpu_spec.py tests/test_gpu_model.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2r_pytest.log
