#!/usr/bin/env python
"""Source-line hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass
--kernel-name regex:<kernel>`: share of the executed warp instructions and of the warp-stall samples per source
line, its global L1 tag requests and its largest stall reason.

usage: python tools/ncu_source_hotspots.py <source.csv> [lines] > profiles/<name>.md"""
import csv
import sys

STALLS = ("stall_long_sb", "stall_math", "stall_wait", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_lg",
          "stall_mio", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_barrier")


def fl(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    out, hdr, fname = [], None, ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
        elif r and r[0].isdigit() and hdr:
            st = {k: fl(r[hdr[k]]) for k in STALLS if k in hdr}
            out.append((fname, int(r[0]), r[1].strip(), fl(r[hdr["Instructions Executed"]]), fl(r[hdr["# Samples"]]),
                        fl(r[hdr["L1 Tag Requests Global"]]), st))
    ti, ts = sum(o[3] for o in out), sum(o[4] for o in out)
    agg = {}
    for o in out:
        for k, v in o[6].items():
            agg[k] = agg.get(k, 0.0) + v
    print(f"warp instructions executed: {ti / 1e6:.1f} M, warp-stall samples: {ts:.0f}\n")
    print("stall samples by reason: " + ", ".join(f"{k[6:]} {100 * v / ts:.1f} %" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0) + "\n")
    print("| file:line | % of warp instructions | % of stall samples | L1 tag requests (M) | largest stall | source |")
    print("|---|---|---|---|---|---|")
    for f, ln, src, inst, samp, l1, st in sorted(out, key=lambda o: -o[4])[:top]:
        reason = max(st.items(), key=lambda kv: kv[1])[0][6:] if st else "-"
        print(f"| {f}:{ln} | {100 * inst / ti:.1f} | {100 * samp / ts:.1f} | {l1 / 1e6:.1f} | {reason} | `{src[:120].replace('|', '/')}` |")


if __name__ == "__main__":
    main()
