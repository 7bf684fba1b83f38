"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    elif unit in ("s", "second"):
        v *= 1e9
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {t / 1e6:10.3f} {t / n / 1e3:10.1f} {100 * t / total:6.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in tot.values()):8d} {total / 1e6:10.3f}")
