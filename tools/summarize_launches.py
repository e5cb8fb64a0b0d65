"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total us, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'n':>6s} {'total us':>12s} {'share':>7s} {'avg us':>9s}")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:6d} {us:12.1f} {us / total:7.1%} {us / n:9.1f}")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):6d} {total:12.1f}")
