#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --prompt 2048 --steps 128 --warmup 8 > gpurun_out/l_bench_8b_p2048.json 2> gpurun_out/l_bench_8b_p2048.err
python -c "
import json; d=json.load(open('gpurun_out/l_bench_8b_p2048.json')); print('decode', d['value'], 'prefill', d['prefill']['value'], 'e2e', d['e2e']['value']); print(d['roofline']); print(d.get('cpu_baseline'))"
tail -3 gpurun_out/l_bench_8b_p2048.err
# ncu --set full of the mat-vec launch sites of one layer + lm_head (4-layer slice: launches 13..17 = layer 3 qkv,o,gu,down + lm_head)
ncu --set full --clock-control none --import-source on -k regex:ps_k_rw_matvec -s 12 -c 5 -f -o gpurun_out/l_rw_matvec python tools/prof_decode.py llama-3.1-8b 4 64 1 > gpurun_out/l_ncu.log 2>&1
tail -2 gpurun_out/l_ncu.log
