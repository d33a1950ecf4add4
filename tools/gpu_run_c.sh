#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_decode.py -x -q ) > gpurun_out/c_pytest.log 2>&1
tail -15 gpurun_out/c_pytest.log
timeout 600 python bench.py --prompt 128 --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/c_bench_8b_p128.json 2> gpurun_out/c_bench_8b_p128.err
cat gpurun_out/c_bench_8b_p128.json; tail -3 gpurun_out/c_bench_8b_p128.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c_launches_8b.csv python bench.py --prompt 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_bench.log 2>&1
python tools/launch_list.py gpurun_out/c_launches_8b.csv > gpurun_out/c_launches_8b.txt 2>&1
head -12 gpurun_out/c_launches_8b.txt
