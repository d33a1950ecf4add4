#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2c_pytest.log
timeout 600 python tools/step_timeline.py llama-3.1-8b 8 2048 > gpurun_out/r2c_step_tl_8b_ctx2048.txt 2>&1; cat gpurun_out/r2c_step_tl_8b_ctx2048.txt
timeout 600 python tools/step_timeline.py llama-3.1-8b 8 64 > gpurun_out/r2c_step_tl_8b_ctx64.txt 2>&1; tail -12 gpurun_out/r2c_step_tl_8b_ctx64.txt
timeout 900 python bench.py --steps 64 --no-cpu-baseline > gpurun_out/r2c_bench_8b.json 2> gpurun_out/r2c_bench_8b.err; cut -c1-400 gpurun_out/r2c_bench_8b.json; tail -3 gpurun_out/r2c_bench_8b.err
