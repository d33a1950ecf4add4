#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1
tail -12 gpurun_out/e_pytest.log
python tools/timeline.py llama-3.1-8b 8 64 > gpurun_out/e_timeline_8b_ctx64.txt 2>&1
python tools/timeline.py llama-3.1-8b 8 2048 > gpurun_out/e_timeline_8b_ctx2048.txt 2>&1
head -18 gpurun_out/e_timeline_8b_ctx64.txt; tail -12 gpurun_out/e_timeline_8b_ctx64.txt; tail -12 gpurun_out/e_timeline_8b_ctx2048.txt
