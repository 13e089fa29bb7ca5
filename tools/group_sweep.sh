#!/bin/bash
# usage (GPU box): bash tools/group_sweep.sh "<group:lanes> ..." — HBM-resident and host-buffer throughput per setting
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for gl in $1; do
  g=${gl%%:*}; l=${gl##*:}
  python bench.py --no-cpu-baseline --no-iso --steps 3 --group $g --lanes $l --distinct 64 > gpurun_out/sweep_$g_$l.json 2> gpurun_out/sweep.err || { echo "$gl failed: $(tail -1 gpurun_out/sweep.err)" >> gpurun_out/sweep.txt; continue; }
  python - "$gl" gpurun_out/sweep_$g_$l.json >> gpurun_out/sweep.txt <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
st = d["stages"]
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "bf_us", st["blackfilter"]["us_per_page"], "deskew_us", st["deskew"]["us_per_page"])
PY
done
cat gpurun_out/sweep.txt
