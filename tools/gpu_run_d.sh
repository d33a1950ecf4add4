#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_decode.py -x -q ) > gpurun_out/d_pytest.log 2>&1
tail -4 gpurun_out/d_pytest.log
python tools/timeline.py llama-3.1-8b 8 64 > gpurun_out/d_timeline_8b_ctx64.txt 2>&1
python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/d_timeline_8b_ctx2048.txt 2>&1
python tools/timeline.py llama-3.1-8b 8 64 --nopdl --nograph > gpurun_out/d_timeline_8b_ctx64_nopdl.txt 2>&1
cat gpurun_out/d_timeline_8b_ctx64.txt; tail -12 gpurun_out/d_timeline_8b_ctx2048.txt;  tail -12 gpurun_out/d_timeline_8b_ctx64_nopdl.txt
