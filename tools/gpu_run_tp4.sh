#!/bin/bash
# 4-GPU pass: TP4 bench line (gpurun --gpus 4 -- 'bash tools/gpu_run_tp4.sh')
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 32 --warmup 4 > gpurun_out/r2z_tp4_bench.json 2> gpurun_out/r2z_tp4_bench.err; tail -2 gpurun_out/r2z_tp4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_tp4_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','prefill','parity','tp')})
PY
