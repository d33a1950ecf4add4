#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cabi_misc.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2t_pytest.log
