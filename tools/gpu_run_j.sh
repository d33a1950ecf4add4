#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 64 > gpurun_out/r2j_bench_8b.json 2> gpurun_out/r2j_bench_8b.err; tail -3 gpurun_out/r2j_bench_8b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_8b.json'))
for k in ('value','ms_per_step','e2e','prefill','extras','e2e_powerserve_stack','parity','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:1200])
print(d['roofline']['step'], d['roofline']['step_incl_kv'])
PY
