"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
per kernel: launches, total us, share of the captured time and (when present) DRAM bytes per launch."""
import csv, re, sys
from collections import defaultdict
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
tot = defaultdict(float); cnt = defaultdict(int); rd = defaultdict(float); wr = defaultdict(float)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^.*::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "")
    m = row.get("Metric Name")
    if m == "gpu__time_duration.sum":
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        tot[name] += us; cnt[name] += 1
    elif m == "dram__bytes_read.sum":
        rd[name] += v * scale.get(unit, 1.0)
    elif m == "dram__bytes_write.sum":
        wr[name] += v * scale.get(unit, 1.0)
total = sum(tot.values())
has_dram = bool(rd) or bool(wr)
hdr = f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}"
if has_dram:
    hdr += f" {'dram_rd_MB/launch':>18s} {'dram_wr_MB/launch':>18s} {'dram_GB/s':>10s}"
print(hdr)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    line = f"{k[:60]:60s} {cnt[k]:8d} {v:12.1f} {v/cnt[k]:10.2f} {100*v/total:6.1f}%"
    if has_dram:
        line += f" {rd[k]/cnt[k]/1e6:18.2f} {wr[k]/cnt[k]/1e6:18.2f} {(rd[k]+wr[k])/(v*1e-6)/1e9:10.1f}"
    print(line)
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {total:12.1f}")
