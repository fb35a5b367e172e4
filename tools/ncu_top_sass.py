"""Top stall-sample SASS lines of an `ncu -i X.ncu-rep --page source --csv --print-source sass` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and r:
        cur["rows"].append(r)
for s in secs:
    h = s["hdr"]; ia = h.index("Warp Stall Sampling (All Samples)"); isrc = h.index("Source"); iex = h.index("Instructions Executed")
    tot = sum(int(r[ia]) for r in s["rows"])
    print("===", s["name"][:90], "samples", tot)
    top = sorted(enumerate(s["rows"]), key=lambda kv: -int(kv[1][ia]))[:topn]
    for idx, r in sorted(top):
        reasons = [f"{k.replace('stall_', '')}={v}" for k, v in zip(h, r) if k.startswith("stall_") and "Not Issued" not in k and v not in ("0", "")]
        print(f"{idx:5d} {int(r[ia]):6d} {100 * int(r[ia]) / max(tot, 1):5.1f}%  ex={r[iex]:>8s}  {r[isrc].strip()[:70]:70s} {' '.join(reasons)[:60]}")
