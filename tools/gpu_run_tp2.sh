#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2z_tp_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r2z_tp2_bench.json 2> gpurun_out/r2z_tp2_bench.err; tail -3 gpurun_out/r2z_tp2_bench.err; head -c 400 gpurun_out/r2z_tp2_bench.json; echo; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_tp2_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','prefill','parity','tp')})
PY
