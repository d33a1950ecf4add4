"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel/grid count, total, mean, share."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    k = row["Kernel Name"].split("(")[0] + " grid=" + row["Grid Size"]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:70]:70s} {v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:9.2f} {v[1] / tot:6.3f}")
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
