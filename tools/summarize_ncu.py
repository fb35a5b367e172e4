"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total us, share."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^.*::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[name] += us; cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {v:12.1f} {v/cnt[k]:10.2f} {100*v/total:6.1f}%")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {total:12.1f}")
