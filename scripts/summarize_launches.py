"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit == "ns":
        v /= 1e3
    elif unit == "ms":
        v *= 1e3
    elif unit == "s":
        v *= 1e6
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total/1e3:.3f} ms over {sum(cnt.values())} launches")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {v:10.1f} {v/cnt[k]:8.1f} {100*v/total:6.1f}%")
