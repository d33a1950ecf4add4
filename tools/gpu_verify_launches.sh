#!/bin/bash
# Launch list of one 12-node tree-verify batch (4 layers of the 8B shape, ctx 2048): gpurun --timeout 900 -- 'bash tools/gpu_verify_launches.sh'
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/verify_launches.csv python tools/tree_verify_timing.py llama-3.1-8b 4 2048 12 > gpurun_out/verify_ncu.log 2>&1
python - <<'PY'
import csv, collections
lines = [l for l in open('gpurun_out/verify_launches.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
# keep the launches of the LAST verify batch: everything after the last ps_k_get_embedding with grid 12
last = max(i for i, r in enumerate(rows) if r['Kernel Name'].startswith('ps_k_get_embedding') and r['Grid Size'].startswith('(12,'))
agg = collections.OrderedDict()
tot = 0
for r in rows[last:]:
    v = float(r['Metric Value'].replace(',', '')); v = v / 1000 if r['Metric Unit'] == 'ns' else v
    k = r['Kernel Name'].split('(')[0] + ' grid=' + r['Grid Size']
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
with open('gpurun_out/verify_launches.txt', 'w') as f:
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k[:72]:72s} {n:4d} {t:9.1f} {t / n:8.2f} {t / tot:6.3f}\n")
    f.write(f"total {tot:.1f} us\n")
print(open('gpurun_out/verify_launches.txt').read())
PY
