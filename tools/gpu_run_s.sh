#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2s_pytest.log
timeout 1500 python bench.py > gpurun_out/r2s_bench_8b.json 2> gpurun_out/r2s_bench_8b.err; tail -3 gpurun_out/r2s_bench_8b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_8b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in d})
print(d.get('roofline',{}).get('step'), d.get('prefill'))
for k,v in d.get('extras',{}).items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','ms_per_pass','roofline_step','why')})
print(d.get('e2e_powerserve_stack')); print(d.get('parity')); print(d.get('cpu_baseline'))
PY
