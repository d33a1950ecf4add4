#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_pytest.log 2>&1
tail -5 gpurun_out/g_pytest.log
timeout 900 python bench.py --prompt 2048 --steps 128 --warmup 8 > gpurun_out/g_bench_8b_p2048.json 2> gpurun_out/g_bench_8b_p2048.err
cat gpurun_out/g_bench_8b_p2048.json; tail -3 gpurun_out/g_bench_8b_p2048.err
timeout 600 python bench.py --model llama-3.2-1b --prompt 128 --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/g_bench_1b.json 2> gpurun_out/g_bench_1b.err
cat gpurun_out/g_bench_1b.json
