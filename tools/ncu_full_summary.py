"""One line per kernel from `ncu --set full` captures: duration, DRAM bytes, achieved GB/s against the measured HBM peak
and the counters that say what bounds the kernel.   python tools/ncu_full_summary.py a.ncu-rep b.ncu-rep ..."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6547.2
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"{rep}: no data"); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:3]:
        def val(n):
            i = hdr.index(n)
            return float(r[i].replace(",", "")) * sc.get(units[i], 1)
        def pct(n):
            return float(r[hdr.index(n)]) if n in hdr else float("nan")
        tu = units[hdr.index("gpu__time_duration.sum")]
        t = val("gpu__time_duration.sum")
        t_us = t * 1000 if tu in ("ms", "msecond") else (t / 1000 if tu in ("ns", "nsecond") else t)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]
        gbs = (rd + wr) / t_us / 1e3
        print(f"{name:30s} {t_us:9.1f} us | dram rd {rd / 1e6:8.2f} MB wr {wr / 1e6:8.2f} MB -> {gbs:7.1f} GB/s = {100 * gbs / peak:5.1f}% of measured HBM peak"
              f" | ncu dram% {pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} tensor% {pct('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in hdr else float('nan'):5.1f}"
              f" issue% {pct('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} warps% {pct('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}"
              f" L1% {pct('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} L2% {pct('lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}"
              f" XU% {pct('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):5.1f}"
              f" | regs {r[hdr.index('launch__registers_per_thread')]} grid {r[hdr.index('launch__grid_size')]} x {r[hdr.index('launch__block_size')]}")
