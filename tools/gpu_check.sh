#!/bin/bash
# One GPU pass of everything the round-end driver runs: gpurun --timeout 3000 -- 'bash tools/gpu_check.sh'
# (tools/gpu_profile.sh: launch list + ncu --set full; tools/gpu_run_tp2.sh: the 2-GPU tensor-parallel pass, gpurun --gpus 2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/check_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/check_smoke.log
timeout 1500 python bench.py > gpurun_out/check_bench_8b.json 2> gpurun_out/check_bench_8b.err; tail -2 gpurun_out/check_bench_8b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/check_bench_8b.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches', 'prefill')})
print('roofline.step', d['roofline'].get('step'))
for k, v in d.get('extras', {}).items():
    print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'ms_per_pass', 'why')})
print(d.get('e2e_powerserve_stack', {}).get('value'), d.get('e2e_powerserve_stack_device_topk', {}).get('value'), d.get('parity', {}).get('logits_bit_exact'))
PY
