"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
print(f"total {total/1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised)")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---|---|---|---|")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name[:110]} | {n} | {ns/1e6:.2f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |")
