#!/bin/bash
# round-2 pass A: parity suite with the new long-context / C-ABI tests, bench (both arms, reduced reference prompt)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 64 > gpurun_out/r2a_bench_8b.json 2> gpurun_out/r2a_bench_8b.err; cat gpurun_out/r2a_bench_8b.json | cut -c1-1500; tail -2 gpurun_out/r2a_bench_8b.err
timeout 600 python bench.py --impl reference --prompt 256 --steps 8 --warmup 2 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; cat gpurun_out/r2a_bench_ref.json | cut -c1-1200; tail -2 gpurun_out/r2a_bench_ref.err
