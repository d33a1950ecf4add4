#!/bin/bash
# 2-GPU pass: tensor-parallel parity tests (batched prefill + decode, all exchange modes) and the TP2 bench line
mkdir -p gpurun_out
echo skip tests
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 > gpurun_out/r2_tp2_bench.json 2> gpurun_out/r2_tp2_bench.err
grep "^{" gpurun_out/r2_tp2_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TP2 decode', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'prefill', d['prefill']['value'], d.get('tp'), d.get('parity'))"
grep -i "error" -A5 gpurun_out/r2_tp2_bench.err | head -20
