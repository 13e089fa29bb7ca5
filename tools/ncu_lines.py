#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel.
usage: ncu_lines.py <nvdisasm -g -c listing> <ncu --page source --csv export> <mangled-name substring> <source file> [top]"""
import collections
import csv
import re
import sys

dis, src_csv, fn, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
m, infn, line = {}, False, None
for l in open(dis):
    if l.lstrip().startswith(".section"):
        infn = (".text." in l) and (fn in l)
        continue
    if not infn:
        continue
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        line = int(mm.group(2))
        continue
    mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if mm:
        m[int(mm.group(1), 16)] = line
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iA, iI, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    if r[0] == "Kernel Name":
        break
    if r[0] != "Address":
        data.append(r)
base = int(data[0][iA], 16)
inst, samp = collections.Counter(), collections.Counter()
for r in data:
    ln = m.get(int(r[iA], 16) - base)
    inst[ln] += int(r[iI])
    samp[ln] += int(r[iS])
ti, ts = sum(inst.values()), sum(samp.values())
src = open(srcfile).read().split("\n")
print(f"total warp instructions {ti}, samples {ts}")
for ln, n in samp.most_common(top):
    print(f"{ln}: samples {100 * n / ts:5.1f}%  inst {100 * inst[ln] / ti:5.1f}%  {src[ln - 1].strip()[:110] if ln else ''}")
