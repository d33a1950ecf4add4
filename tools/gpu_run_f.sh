#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_decode.py -x -q ) > gpurun_out/f_pytest.log 2>&1
tail -5 gpurun_out/f_pytest.log
python tools/timeline.py llama-3.1-8b 8 64 > gpurun_out/f_timeline_8b_ctx64.txt 2>&1
head -9 gpurun_out/f_timeline_8b_ctx64.txt; tail -12 gpurun_out/f_timeline_8b_ctx64.txt
