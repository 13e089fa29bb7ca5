#!/usr/bin/env python
"""Turn one `ncu --set full` capture of ONE group (bench.py --pages G --group G --lanes 1) into
  * a markdown table of every kernel launch (time, DRAM bytes, instructions, issue/warp activity) and
  * profiles/traffic.json: DRAM bytes per sheet of every pipeline stage (dram__bytes_read.sum +
    dram__bytes_write.sum of the stage's launches / sheets per launch).

usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_stage_traffic.py <label> <sheets per launch> <source text> [traffic.json]

Launches are assigned to stages by walking the engine's launch order (csrc/engine.c:issue_group)."""
import csv
import json
import sys

STAGES = [
    ("decode", ["k_page_reset", "k_fill_jobs", "k_copy_jobs"]),
    ("blackfilter", ["k_zero_u32", "k_linesum_cols", "k_linesum_cols4", "k_linesum_cols16", "k_linesum_rows", "k_bf_scan"]),
    ("noisefilter", ["k_nf_classify_bits", "k_nf_classify_g8", "k_nf_classify", "k_nf_resolve"]),
    ("blurfilter", ["k_rect_count", "k_blur_decide_sm", "k_blur_decide", "k_blur_wipe"]),
    ("grayfilter", ["k_cellstats", "k_zero_range", "k_gray_prewhite", "k_gray_windows", "k_gray_cascade", "k_gray_wipe"]),
    ("detect_masks", ["k_zero_u32", "k_linesum_cols", "k_linesum_cols4", "k_linesum_cols16", "k_linesum_rows", "k_detect_edges", "k_assemble_masks2"]),
    ("detect_rotation", ["k_rot_colprefix", "k_rot_peaks_w", "k_rot_peaks_h", "k_rot_peaks", "k_rot_finalize", "k_rot_set"]),
    ("deskew", ["k_inkmap", "k_rotate_sheet", "k_rotate_sheet_g8c", "k_swap_sheets"]),
    ("center_mask", ["k_zero_u32", "k_linesum_cols", "k_linesum_cols4", "k_linesum_cols16", "k_linesum_rows", "k_detect_edges", "k_assemble_masks2", "k_prep_center_move",
                     "k_move_pass", "k_swap_sheets"]),
    ("border", ["k_zero_u32", "k_linesum_cols", "k_linesum_cols4", "k_linesum_cols16", "k_linesum_rows", "k_detect_border", "k_border_to_mask", "k_prep_align_move",
                "k_prep_border_maskjob", "k_apply_masks", "k_set_other", "k_move_pass", "k_swap_sheets"]),
    ("output", ["k_pack_rows", "k_convert_out"]),
]


def main():
    label, sheets, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out_json = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]]
        return v * {"ms": 1e3, "msecond": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3, "s": 1e6, "second": 1e6,
                    "Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)

    # the capture may start in the middle of a group and run into the next one: rotate it so that it
    # starts at the group's first launch (k_page_reset) and keep one full group
    body = rows[2:]
    first = next((i for i, r in enumerate(body) if r[col["Kernel Name"]].startswith("k_page_reset")), 0)
    head = body[:first]                      # the tail of the previous group
    body = body[first:]
    nxt = next((i for i, r in enumerate(body[1:], 1) if r[col["Kernel Name"]].startswith("k_page_reset")), len(body))
    body = body[:nxt]
    # what the capture did not reach of this group (e.g. the output stage) is taken from the previous group's tail
    names_main = [r[col["Kernel Name"]].split("(")[0] for r in body]
    last_stage = max((i for i, (_, ks) in enumerate(STAGES) if any(k in names_main for k in ks[-1:])), default=-1)
    for r in head:
        nm = r[col["Kernel Name"]].split("(")[0]
        st = max((i for i, (_, ks) in enumerate(STAGES) if nm in ks), default=-1)
        if st > last_stage and all(nm not in ks for _, ks in STAGES[:last_stage + 1]):
            body.append(r)
    si, seen_in_stage = 0, False
    per_stage = {}
    print(f"| capture | stage | kernel | time (us) | per sheet (us) | DRAM read (MB) | DRAM write (MB) | traffic / sheet (MB) | "
          f"warp inst (M) | issue active % | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0]
        # advance to the first stage (from the current one on) that knows this kernel; a kernel that
        # belongs to the current stage keeps it
        j = si
        while j < len(STAGES) and name not in STAGES[j][1]:
            j += 1
        if j == len(STAGES):
            j = 0             # the capture ran into the next group: start over at decode
            while j < len(STAGES) and name not in STAGES[j][1]:
                j += 1
            if j == len(STAGES):
                continue
        if j != si:
            si = j
        # the first kernel of a later stage that also exists in the current one (k_zero_u32 ...) starts the next stage
        elif name == STAGES[si][1][0] and seen_in_stage and si + 1 < len(STAGES) and any(name == s[1][0] for s in STAGES[si + 1:si + 2]):
            si += 1
        seen_in_stage = True
        stage = STAGES[si][0]
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        inst = val(r, "smsp__inst_executed.sum") / 1e6
        issue = float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
        warps = float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])
        a = per_stage.setdefault(stage, {"us": 0.0, "mb": 0.0, "kernels": []})
        a["us"] += t
        a["mb"] += rd + wr
        a["kernels"].append(name)
        print(f"| {label} | {stage} | {name} | {t:.1f} | {t / sheets:.2f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / sheets:.2f} | "
              f"{inst:.1f} | {issue:.1f} | {warps:.1f} |")
    print()
    print("| stage | kernel time per sheet (us, serialised, cold) | DRAM traffic per sheet (MB) |")
    print("|---|---|---|")
    tj = {}
    for stage, _ in STAGES:
        if stage in per_stage:
            a = per_stage[stage]
            print(f"| {stage} | {a['us'] / sheets:.2f} | {a['mb'] / sheets:.2f} |")
            tj[stage] = {"bytes_per_sheet": round(a["mb"] * 1e6 / sheets), "source": source,
                         "kernels": sorted(set(a["kernels"]))}
    if out_json:
        with open(out_json, "w") as f:
            json.dump(tj, f, indent=1)


if __name__ == "__main__":
    main()
