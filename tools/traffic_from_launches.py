"""DRAM bytes per launch of the tensor-core kernel family (tc_gemm*, attn_fwd) from an ncu launch list with
dram__bytes_read.sum / dram__bytes_write.sum -> profiles/rN_traffic.json (read by bench.py for `roofline.traffic`).
    python tools/traffic_from_launches.py gpurun_out/c_launches_step.csv profiles/r2_traffic.json "<source description>" """
import csv, json, re, sys
from collections import defaultdict
src, dst, desc = sys.argv[1], sys.argv[2], sys.argv[3]
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = defaultdict(lambda: defaultdict(float))
for row in csv.DictReader(lines):
    name = re.sub(r"^.*::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
    v = float(row["Metric Value"].replace(",", "")); unit = row.get("Metric Unit", "")
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        per[row["ID"]]["us"] = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    else:
        per[row["ID"]][m] = v * scale.get(unit, 1.0)
    per[row["ID"]]["name"] = name
fam = [d for d in per.values() if str(d["name"]).startswith(("tc_gemm", "attn_fwd"))]
tot_us = sum(d["us"] for d in per.values())
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in fam); wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in fam)
out = {"source": desc, "kernel": "tc_gemm_kernel / tc_gemm2_kernel / tc_gemm_swap_kernel / attn_fwd_kernel (all template instances)",
       "launches_per_step": len(fam), "dram_bytes_per_launch": (rd + wr) / len(fam), "dram_read_bytes_per_launch": rd / len(fam),
       "dram_write_bytes_per_launch": wr / len(fam), "share_of_step_time": sum(d["us"] for d in fam) / tot_us,
       "avg_launch_us_under_ncu": sum(d["us"] for d in fam) / len(fam)}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
