#!/bin/bash
# round-1 final single-GPU pass: smoke, parity suite, both bench arms, 1B config
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/q_bench_8b.json 2> gpurun_out/q_bench_8b.err; cat gpurun_out/q_bench_8b.json; tail -2 gpurun_out/q_bench_8b.err
timeout 600 python bench.py --model llama-3.2-1b --prompt 128 --steps 256 > gpurun_out/q_bench_1b.json 2> gpurun_out/q_bench_1b.err; cat gpurun_out/q_bench_1b.json; tail -2 gpurun_out/q_bench_1b.err
