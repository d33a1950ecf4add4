#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_tp.py -x -q -k "slice-1b-4" ) > gpurun_out/tp4_pytest.log 2>&1
grep -v "^$" gpurun_out/tp4_pytest.log | tail -8 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --prompt 256 --steps 64 --warmup 8 > gpurun_out/tp4_bench.json 2> gpurun_out/tp4_bench.err
grep "^{" gpurun_out/tp4_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TP4 decode', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('tp'), d['config']['parallelism'], d['scaling'])"
grep -i "error" -A5 gpurun_out/tp4_bench.err | head -20
