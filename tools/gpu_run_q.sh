#!/bin/bash
# round-1 final single-GPU measurement pass: parity suite, both bench arms, launch list + ncu --set full of the mat-vec / attention kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/q_bench_8b.json 2> gpurun_out/q_bench_8b.err; cat gpurun_out/q_bench_8b.json; tail -2 gpurun_out/q_bench_8b.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/q_bench_8b_reference.json 2> gpurun_out/q_bench_8b_reference.err; cat gpurun_out/q_bench_8b_reference.json; tail -2 gpurun_out/q_bench_8b_reference.err
timeout 600 python bench.py --model llama-3.2-1b --prompt 128 --steps 256 --no-cpu-baseline > gpurun_out/q_bench_1b.json 2> gpurun_out/q_bench_1b.err; cat gpurun_out/q_bench_1b.json; tail -2 gpurun_out/q_bench_1b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/q_launches_8b.csv python bench.py --prompt 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/q_ncu_bench.log 2>&1
python tools/launch_list.py gpurun_out/q_launches_8b.csv > gpurun_out/q_launches_8b.txt 2>&1; head -14 gpurun_out/q_launches_8b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ps_k_rw_matvec|ps_k_attn1|ps_k_attn2" -s 18 -c 7 -f -o gpurun_out/q_decode_kernels python tools/prof_decode.py llama-3.1-8b 4 2048 1 > gpurun_out/q_ncu.log 2>&1
tail -2 gpurun_out/q_ncu.log
