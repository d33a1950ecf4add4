#!/bin/bash
# first GPU pass of the session: parity tests, decode bench (1B, 8B), launch list, per-CTA trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
python bench.py --model llama-3.2-1b --prompt 128 --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/a_bench_1b.json 2> gpurun_out/a_bench_1b.err
python bench.py --prompt 128 --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/a_bench_8b_p128.json 2> gpurun_out/a_bench_8b_p128.err
timeout 900 python bench.py --prompt 2048 --steps 128 --warmup 8 > gpurun_out/a_bench_8b_p2048.json 2> gpurun_out/a_bench_8b_p2048.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/a_launches_8b.csv python bench.py --prompt 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_bench.log 2>&1
python tools/launch_list.py gpurun_out/a_launches_8b.csv > gpurun_out/a_launches_8b.txt 2>&1
python tools/trace_matvec.py llama-3.1-8b 4 > gpurun_out/a_trace_8b.txt 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_bench_*.json; head -20 gpurun_out/a_launches_8b.txt; cat gpurun_out/a_trace_8b.txt
