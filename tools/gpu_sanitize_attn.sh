#!/bin/bash
# New tests + compute-sanitizer (memcheck, racecheck) over the tiled attention kernels: gpurun --timeout 1500 -- 'bash tools/gpu_sanitize_attn.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k tiled 2>&1 | tail -4 | tee gpurun_out/san_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiled and hs128 and 76" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/san_memcheck.log; tail -3 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "tiled and r7 and 100" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/san_racecheck.log
