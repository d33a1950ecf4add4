#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/tp2_bench.json 2> gpurun_out/tp2_bench.err
grep "^{" gpurun_out/tp2_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TP2 decode', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('tp'), d['config']['parallelism'], d['scaling'])"
grep -i "error" -A5 gpurun_out/tp2_bench.err | head -20
