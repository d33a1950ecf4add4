#!/bin/bash
# per-op executor dispatch (a14): op tests + the unfused PowerServe graph on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cabi_misc.py tests/test_gpu_dropin.py tests/test_gpu_ops.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2o_pytest.log
