"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel
launches, average and total duration, share.  usage: python tools/launch_shares.py file.csv [skip_launches]"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v / 1000.0 if unit in ("ns", "nsecond") else v * (1000.0 if unit in ("ms", "msecond") else 1.0)
    rows.append((r["Kernel Name"].split("(")[0], v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1
    agg[k][1] += v
print(f"{len(rows)} launches, {tot / 1000.0:.2f} ms")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {n:5d} launches  avg {v / n:9.1f} us  total {v / 1000.0:8.2f} ms  {100.0 * v / tot:5.1f} %")
