#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into rows of profiles/r01_ncu_summary.md.

usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv | python tools/ncu_rows.py <label> <sheets per launch>
Only the first launch of every kernel name is listed (the capture holds each twice)."""
import csv
import sys

label, sheets = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}


def val(r, name, want_unit=None):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"ms": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "s": 1e6,
             "Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
    return v * scale


seen = set()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    inst = val(r, "smsp__inst_executed.sum") / 1e6
    issue = val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
    warps = val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    print(f"| {label} | {name} | {t:.1f} | {t / sheets:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / sheets:.2f} | "
          f"{inst:.1f} | {issue:.1f} | {warps:.1f} |")
