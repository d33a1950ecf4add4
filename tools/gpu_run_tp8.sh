#!/bin/bash
# N-GPU pass of the bench exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 64 --warmup 8 > gpurun_out/r2_tp${N}_bench.json 2> gpurun_out/r2_tp${N}_bench.err
grep "^{" gpurun_out/r2_tp${N}_bench.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('TP', d['n_gpus'], 'decode', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'prefill', d['prefill']['value'], d.get('tp'), d.get('parity'))"
grep -i "error\|Traceback" -A8 gpurun_out/r2_tp${N}_bench.err | head -30
