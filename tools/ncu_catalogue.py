#!/usr/bin/env python
"""Kernel catalogue from the raw page of one `ncu --set full` capture: launch geometry, registers, shared memory,
what the launch did (time, DRAM bytes, warp instructions) and what kept it from going faster (issue / pipe / LSU
activity, cache hit rates, the two largest warp stall reasons).

usage: python tools/ncu_catalogue.py <raw.csv> <sheets per launch> > profiles/<name>.md"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    sheets = int(sys.argv[2])
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def f(r, name, default=0.0):
        try:
            return float(r[col[name]].replace(",", ""))
        except (KeyError, ValueError):
            return default

    def scaled(r, name):
        u = units[col[name]]
        return f(r, name) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3,
                             "byte": 1e-6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)

    stalls = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    print("| # | kernel | grid | block | regs | smem KB (static + dynamic) | us | us / sheet | DRAM MB / sheet | warp inst M | issue % | "
          "FMA pipe % | ALU pipe % | LSU wavefronts % | L1 hit % | L2 hit % | warps active % | largest stall reasons (warps per issue) |")
    print("|" + "---|" * 18)
    for i, r in enumerate(body):
        name = r[col["Kernel Name"]].split("(")[0]
        st = sorted(((f(r, n), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for n in stalls
                     if "selected" not in n), reverse=True)[:2]
        smem = f"{f(r, 'launch__shared_mem_per_block_static'):.1f} + {f(r, 'launch__shared_mem_per_block_dynamic'):.1f}"
        t = scaled(r, "gpu__time_duration.sum")
        mb = scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")
        print(f"| {i} | {name} | {r[col['Grid Size']]} | {r[col['Block Size']]} | {int(f(r, 'launch__registers_per_thread'))} | {smem} | "
              f"{t:.1f} | {t / sheets:.2f} | {mb / sheets:.2f} | {f(r, 'smsp__inst_executed.sum') / 1e6:.1f} | "
              f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{f(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{f(r, 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{f(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} | "
              f"{f(r, 'l1tex__t_sector_hit_rate.pct'):.0f} | {f(r, 'lts__t_sector_hit_rate.pct'):.0f} | "
              f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
              f"{st[0][1]} {st[0][0]:.1f}, {st[1][1]} {st[1][0]:.1f} |")


if __name__ == "__main__":
    main()
