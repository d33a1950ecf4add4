#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/h_pytest.log 2>&1
tail -6 gpurun_out/h_pytest.log
timeout 900 python bench.py --prompt 2048 --steps 128 --warmup 8 --no-cpu-baseline > gpurun_out/h_bench_8b_p2048.json 2> gpurun_out/h_bench_8b_p2048.err
python -c "
import json; d=json.load(open('gpurun_out/h_bench_8b_p2048.json')); print('decode', d['value'], 'prefill', d['prefill'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/h_bench_8b_p2048.err
