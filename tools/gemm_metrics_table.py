"""Per-launch table of the tensor-core kernels of one step from an ncu CSV (--page raw style log with the metrics below).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\\
lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        -k regex:'tc_gemm|attn_fwd' --csv --log-file launches.csv python bench.py --profile-step
    python tools/gemm_metrics_table.py launches.csv"""
import collections
import csv
import re
import sys


def unit_scale(unit, base):
    u = (unit or "").lower()
    table = {"": 1.0, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
             "msecond": 1e3, "nsecond": 1e-3, "%": 1.0}
    return table.get(u, 1.0)


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, mi, ui, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[ii].isdigit():
            continue
        d = launches.setdefault(r[ii], {"k": r[ki]})
        try:
            d[r[mi]] = float(r[vi].replace(",", "")) * unit_scale(r[ui], None)
        except ValueError:
            pass
    print("  # kernel                               us  tensor%  L2 thr%  L2->SM MB   DRAM MB")
    tot_us = tot_tensor = 0.0
    agg = collections.OrderedDict()
    for n, (i, d) in enumerate(launches.items()):
        name = re.sub(r"^.*?(tc_gemm\w*|attn_fwd)_kernel", r"\1_kernel", d["k"]).split("(")[0]
        us = d.get("gpu__time_duration.sum", 0.0)
        te = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        l2 = d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        xb = d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0.0) / 1e6
        dr = (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6
        print(f"{n:3d} {name:32s} {us:8.1f} {te:8.1f} {l2:8.1f} {xb:10.1f} {dr:9.1f}")
        tot_us += us
        tot_tensor += us * te
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += us * te
    print()
    for name, (c, us, w) in agg.items():
        print(f"{name:32s} {c:3d} launches {us:9.1f} us  time-weighted tensor-pipe active {w / max(us, 1e-9):5.1f} %")
    print(f"{'all tensor-core launches':32s} {len(launches):3d} launches {tot_us:9.1f} us  time-weighted tensor-pipe active {tot_tensor / max(tot_us, 1e-9):5.1f} %")


if __name__ == "__main__":
    main()
