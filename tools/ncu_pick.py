#!/usr/bin/env python
"""Key counters per kernel from an `ncu --page raw --csv` export: duration, DRAM bytes and throughput, issue activity."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = {k: h.index(k) for k in keys if k in h}
iname = h.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[iname]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
    seen.setdefault(name, []).append(r)
for name, rs in seen.items():
    r = rs[-1]  # last launch of the kernel
    print(f"{name}  (launches captured: {len(rs)})")
    for k, i in idx.items():
        print(f"    {k:60s} {r[i]:>16s} {rows[1][i]}")
