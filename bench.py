#!/usr/bin/env python
"""bench.py — A4 300-dpi GRAY8 pages/sec through the full per-sheet pipeline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: one process per GPU (torchrun for N>1; pages are independent, so the
job is sharded with no collective — weak scaling, `pages` sheets per rank per
step).  A step = one pass of the whole process_sheet() pipeline (decode ->
blackfilter -> noisefilter -> blurfilter -> grayfilter -> detect masks ->
detect rotation -> deskew (cubic) -> centre -> border scan/apply/align ->
output) over one batch of synthetic BASELINE-config-2 pages.
`value`: inputs already resident in HBM.  `e2e`: the same pipeline through
unpaper_b200_engine_process_host() with pinned HOST buffers — H2D of every page
and D2H of every finished sheet inside the timed region.

Reference arm (--impl reference): the reference's own CPU backend and its own
process_sheet() (oracle/_ref, compiled from the unmodified sources) on the
box's host cores, all cores, same pages; each step a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware queue per engine lane (the default of 8 makes lanes share queues)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

import unpaper_gpu_b200 as U  # noqa: E402
from unpaper_gpu_b200 import synth  # noqa: E402
from unpaper_gpu_b200 import shard  # noqa: E402
from oracle import checker  # noqa: E402  test infrastructure: the CPU checkers

W, H = synth.A4_W, synth.A4_H
WORKLOAD = ("BASELINE config 2: synthetic A4 300-dpi GRAY8 2480x3508, +-5 deg skew, speckle 1/5000, "
            "dark scan edges; default single-layout pipeline (black/noise/blur/gray filters, mask scan, "
            "deskew cubic, mask centring, border scan+align)")
# BASELINE.json:metric names "A4 300dpi GRAY8 pages/sec full pipeline at 1/2/4/8 B200; per-kernel HBM GB/s":
# the first part is this metric, the per-kernel part is the `roofline` object of the line
METRIC, UNIT = "A4 300dpi GRAY8 pages/sec full pipeline", "pages/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def _one_page(seed):
    return synth.gray_page(seed, W, H)


def make_pages(n_distinct, rank):
    """`n_distinct` pages with per-page seeds (BASELINE config 5), generated on the host
    cores in parallel.  Must run before CUDA is initialised in this process (fork)."""
    seeds = [rank * 100003 + i for i in range(n_distinct)]
    nproc = min(len(seeds), max(1, len(os.sched_getaffinity(0))), 32)
    if nproc <= 1:
        return np.stack([_one_page(s) for s in seeds])
    import multiprocessing as mp
    with mp.get_context("fork").Pool(nproc) as pool:
        return np.stack(pool.map(_one_page, seeds, chunksize=max(1, len(seeds) // (4 * nproc))))


def run_reference(args, rank, world):
    """The reference's CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    lib = checker.load_ref()
    kind = "reference"
    prefix = "ref_"
    if lib is None:
        lib, kind, prefix = checker.load_oracle(), "port", "orc_"
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref nor oracle/liboracle.so is built"}))
        return
    cores = os.cpu_count() or 1
    cfg = U.default_sheet_config()
    sample = cores                       # one page per core per step
    pages = make_pages(min(sample, 8), 0)
    pages = np.concatenate([pages] * ((sample + len(pages) - 1) // len(pages)))[:sample]
    budget_s = 240.0
    t_begin = time.time()
    times = []
    for i in range(min(args.warmup, 1) + args.steps):
        t0 = time.time()
        checker.process_sheets_cpu(lib, prefix, cfg, pages, W, H, U.FMT_GRAY8, threads=cores, want_out=False)
        dt = time.time() - t0
        if i >= min(args.warmup, 1):
            times.append(dt)
        if time.time() - t_begin + dt > budget_s and times:
            break
    ms = 1000.0 * sum(times) / len(times)
    v = sample / (ms / 1000.0)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_step": sample, "threads": cores,
                       "note": "reference CPU backend process_sheet(), pages injected in memory, no codecs; "
                               "step count bounded to ~4 min of wall clock"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sample} pages per step, {len(times)} timed steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# algorithmic (compulsory) bytes per sheet of each stage, S = working-sheet bytes,
# M = mask-rectangle bytes (bpp = 1 for the gray working sheet).
# deskew: SURVEY's "2 S if fused into a fresh sheet" (it is: one sweep into the slot's other buffer);
# center_mask / border: the mask scan's read of the sheet + the one sweep that replaces SURVEY's five
# blits of 5 M (the blits are not compulsory, and counting them would put `alg_gbs` above the HBM peak);
# output: nothing is moved when the last sweep already wrote the caller's buffer.
def stage_bytes(S, M, direct_out=True):
    return {"decode": 2 * S, "blackfilter": S, "noisefilter": S, "blurfilter": S, "grayfilter": S,
            "detect_masks": S, "detect_rotation": 0.25 * S, "deskew": 2 * S, "center_mask": S + 2 * S,
            "border": S + 2 * S, "output": 0 if direct_out else 2 * S}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pages", type=int, default=int(os.environ.get("BENCH_PAGES", "1024")), help="sheets per rank per step (HBM-resident arm)")
    ap.add_argument("--e2e-pages", type=int, default=int(os.environ.get("BENCH_E2E_PAGES", "1024")), help="sheets per rank per step (host-buffer arm)")
    ap.add_argument("--group", type=int, default=int(os.environ.get("BENCH_GROUP", "32")))
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("BENCH_LANES", "8")))
    ap.add_argument("--distinct", type=int, default=int(os.environ.get("BENCH_DISTINCT", "256")),
                    help="distinct synthetic pages per rank (per-page seeds); the step's pages cycle through them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-iso", action="store_true", help="skip the isolated per-stage pass (profiling runs)")
    ap.add_argument("--iso-group", type=int, default=int(os.environ.get("BENCH_ISO_GROUP", "296")),
                    help="sheets per launch sequence of the isolated per-stage pass: 2 x 148 SMs, so that the kernels "
                         "that give one CTA to a page (flood fill, noise resolve, gray cascade) fill the machine")
    ap.add_argument("--out-format", default="page", choices=["page", "mono"],
                    help="sheet_stage_output format: 'page' (GRAY8, the headline config) or 'mono' (pbm, 1 bit/px D2H)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    affinity = shard.bind_near_gpu(local, world) if world > 1 else None   # before CUDA/pinned allocations
    args.distinct = max(1, min(args.distinct, args.pages))
    distinct = make_pages(args.distinct, rank)                      # before CUDA exists in this process (fork)
    import torch
    import torch.distributed as dist
    from unpaper_gpu_b200.lib import Engine

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = U.default_sheet_config()
    eng = Engine(cfg, W, H, U.FMT_GRAY8, group_pages=args.group, lanes=args.lanes, device=local)
    if args.out_format == "mono":
        eng.set_output_format(U.FMT_MONOWHITE)
    out_bytes = eng.sheet_bytes
    reps = (args.pages + args.distinct - 1) // args.distinct
    host_np = np.concatenate([distinct] * reps)[:args.pages]
    e2e_pages = min(args.e2e_pages, args.pages)
    host_in = torch.from_numpy(host_np[:e2e_pages]).pin_memory()
    host_out = torch.empty((e2e_pages, H, W), dtype=torch.uint8).pin_memory()
    dev_in = torch.from_numpy(distinct).to(f"cuda:{local}").repeat((reps, 1, 1))[:args.pages].contiguous()
    dev_out = torch.empty((args.pages, H, W), dtype=torch.uint8, device=f"cuda:{local}")
    res = (U.SheetResult * args.pages)()

    def pcie_gbs():
        """Raw pinned-memory copy bandwidth: the ceiling of the e2e arm.  Every phase starts
        behind a barrier, so with N ranks all N GPUs copy at the same time and the sum of the
        per-rank rates is what the node (not one link) can move."""
        n = min(e2e_pages, 64)
        t_h2d = t_d2h = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(); dev_out[:n].copy_(host_in[:n], non_blocking=True); b.record(); b.synchronize()
            t_h2d = min(t_h2d, a.elapsed_time(b))
            barrier()
            a.record(); host_out[:n].copy_(dev_out[:n], non_blocking=True); b.record(); b.synchronize()
            t_d2h = min(t_d2h, a.elapsed_time(b))
        # both directions at once (what the e2e arm asks of the link): two streams, one clock
        t_bi = 1e9
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        scratch = torch.empty_like(dev_out[:n])
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record()
            s1.wait_event(a); s2.wait_event(a)
            with torch.cuda.stream(s1):
                scratch.copy_(host_in[:n], non_blocking=True)
            with torch.cuda.stream(s2):
                host_out[:n].copy_(dev_out[:n], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            b.record(); b.synchronize()
            t_bi = min(t_bi, a.elapsed_time(b))
        gb = n * W * H / 1e9
        mine = [gb / (t_h2d / 1e3), gb / (t_d2h / 1e3), gb / (t_bi / 1e3)]
        node = shard.sum_over_ranks(mine, device=f"cuda:{local}")       # all ranks copied at the same time
        return [round(x, 1) for x in mine], [round(x, 1) for x in node]

    def step_device():
        eng.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), args.pages, False, res)
        return eng.last_device_ms()

    def step_host():
        eng.process_ptr(host_in.data_ptr(), host_out.data_ptr(), e2e_pages, True, res)
        return eng.last_device_ms()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        eng.set_profiling(profile)
        barrier()
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(steps):
            dev_ms += fn()          # CUDA events on the engine's own streams, summed over steps
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        t = shard.max_over_ranks([dev_ms, wall_ms], device=f"cuda:{local}")
        return t[0], t[1], eng.launch_count() - l0

    (h2d_gbs, d2h_gbs, bidir_gbs), node_gbs = pcie_gbs()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, dev_wall_ms, launches = timed(step_device, args.steps, max(args.warmup, 3), profile=True)
    prof = eng.profile()
    spread = eng.profile_spread()
    eng.set_profiling(False)
    e2e_ms, e2e_wall_ms, _ = timed(step_host, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None

    # correctness guard: every sheet deskewed and flagged ok
    bad = sum(1 for r in res if r.status != 0)
    total_pages = args.pages * world
    value = total_pages * args.steps / (dev_ms / 1000.0)
    e2e_value = e2e_pages * world * args.steps / (e2e_ms / 1000.0)

    # Per-stage durations with nothing else on the GPU: a second, single-lane engine runs
    # two groups with CUDA events around every stage (rank 0 only; the lanes of the main
    # engine are idle by now).  This is "the kernel's own launch duration"; the numbers
    # taken inside the timed region (all lanes competing for the SMs) are reported too.
    iso = None
    if rank == 0 and args.no_iso:
        iso = (prof, args.group)
    elif rank == 0:
        n_iso = max(1, min(args.iso_group, args.pages))
        ie = Engine(cfg, W, H, U.FMT_GRAY8, group_pages=n_iso, lanes=1, device=local)
        for _ in range(2):
            ie.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), n_iso, False, None)
        ie.set_profiling(True)
        for _ in range(3):
            ie.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), n_iso, False, None)
        iso = (ie.profile(), n_iso)
        ie.close()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        S = W * H
        M = S * 0.82        # typical detected mask share of the sheet on these pages
        sb = stage_bytes(S, M, direct_out=(args.out_format == "page"))

        def table(profile, pages_per_group):
            out, dom, dom_ms = {}, None, 0.0
            for k, (ms, cnt) in profile.items():
                if cnt == 0:
                    continue
                avg_ms = ms / cnt                   # one launch sequence = one group of sheets
                gbs = sb.get(k, 0) * pages_per_group / (avg_ms / 1000.0) / 1e9 if avg_ms > 0 else 0.0
                out[k] = {"ms_per_group": round(avg_ms, 4), "alg_gbs": round(gbs, 1),
                          "us_per_page": round(1000.0 * avg_ms / pages_per_group, 2)}
                if ms > dom_ms:
                    dom, dom_ms = k, ms
            return out, dom

        per_stage, dom_conc = table(prof, args.group)
        iso_stage, dom = table(iso[0], iso[1])
        traffic, traffic_src = None, None
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tj):
            with open(tj) as f:
                tr = json.load(f).get(dom)
            if tr:   # dram__bytes_read.sum + dram__bytes_write.sum per sheet, from the committed ncu capture
                traffic, traffic_src = tr["bytes_per_sheet"] * iso[1], tr["source"]
        a = iso_stage[dom]["alg_gbs"]
        roof = {"bound": "hbm", "kernel": dom, "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(a / peaks["hbm_gbs"], 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sb.get(dom, 0) * iso[1],
                "launch": f"one stage launch sequence over a group of {iso[1]} sheets, timed alone with CUDA events "
                          "on the launching stream",
                "concurrent": {"kernel": dom_conc, "achieved": per_stage[dom_conc]["alg_gbs"],
                               "frac": round(per_stage[dom_conc]["alg_gbs"] / peaks["hbm_gbs"], 4),
                               "note": "same stage timed inside the timed region with all lanes competing"},
                "whole_sheet": {"algorithmic_bytes_per_sheet": sum(sb.values()),
                                "achieved": round(sum(sb.values()) * value / world / 1e9, 1), "unit": "GB/s",
                                "frac": round(sum(sb.values()) * value / world / 1e9 / peaks["hbm_gbs"], 4),
                                "note": "all stages: sum of the per-stage algorithmic bytes x measured sheets/s per GPU "
                                        "(the throughput arm, all lanes overlapping)"},
                "note": "algorithmic bytes per SURVEY 8(d) with the 1 B/px working sheet, fused forms (deskew 2 S, mask moves one "
                        "sweep of 2 S each, no output copy); the dominant stage is "
                        "the one with the largest share of the isolated per-sheet time. The stages are not HBM-bound "
                        "in this implementation: the limits are instruction issue (rotate, noise classification) "
                        "and chains of dependent steps (flood fill, cascades) - see DESIGN.md section 5"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "output_format": args.out_format, "pages_per_step_per_gpu": args.pages, "group_pages": args.group,
                           "lanes": args.lanes, "distinct_pages": args.distinct,
                           "l2": f"inputs larger than L2 ({args.pages * S / 1e6:.0f} MB of pages per step vs 126 MB L2)",
                           "parallelism": f"page-sharded x{world}, no collective", "cpu_affinity": affinity,
                           "timing": "CUDA events on the engine's streams (first enqueue -> last lane done), max over ranks",
                           "wall_ms_per_step": dev_wall_ms / args.steps},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_pages * S * world,
                        "d2h_bytes_per_step": e2e_pages * out_bytes * world, "ms_per_step": e2e_ms / args.steps,
                        "wall_ms_per_step": e2e_wall_ms / args.steps, "pages_per_step_per_gpu": e2e_pages,
                        "pcie_h2d_gbs": h2d_gbs, "pcie_d2h_gbs": d2h_gbs,
                        "pcie_bidir_gbs_each_way": bidir_gbs,
                        "pcie_bound_pages_per_sec_per_gpu": round(min(h2d_gbs, d2h_gbs, bidir_gbs) * 1e9 / S, 1),
                        # the node ceiling: every rank copying at once, both directions (sum over ranks, GB/s each way)
                        "node_copy_gbs": {"h2d": node_gbs[0], "d2h": node_gbs[1], "bidir_each_way": node_gbs[2]},
                        "node_copy_bound_pages_per_sec": round(node_gbs[2] * 1e9 / S, 1),
                        "frac_of_node_copy_bound": round(e2e_value / (node_gbs[2] * 1e9 / S), 4)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "stages": per_stage,
                "stages_isolated": iso_stage,
                # fastest / slowest group of the serial-per-page stages inside the timed region (pages differ:
                # a group finishes with its slowest page)
                "group_spread_ms": {k: [round(v[0], 4), round(v[1], 4)] for k, v in spread.items()
                                    if k in ("blackfilter", "noisefilter", "grayfilter", "deskew")},
                "failed_sheets": bad}
        if world == 1 and not args.no_cpu_baseline:
            lib = checker.load_ref()
            kind, prefix = "reference", "ref_"
            if lib is None:
                lib, kind, prefix = checker.load_oracle(), "port", "orc_"
            if lib is not None and hasattr(lib, prefix + "process_sheets"):
                cores = os.cpu_count() or 1
                sample = host_np[:min(cores, args.pages)]
                if len(sample) < cores:
                    sample = np.concatenate([sample] * ((cores + len(sample) - 1) // len(sample)))[:cores]
                t0 = time.time()
                ref_out, _ = checker.process_sheets_cpu(lib, prefix, cfg, sample, W, H, U.FMT_GRAY8, threads=cores, want_out=True)
                dt = time.time() - t0
                line["cpu_baseline"] = {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": f"{len(sample)} pages of the same workload, {cores} threads, one pass ({dt:.1f} s)"}
                # the checker's sheets are at hand: hold the timed e2e output of the same pages to them
                nv = min(len(sample), e2e_pages, args.pages)
                got = host_out[:nv].numpy().reshape(nv, -1)
                bad_bytes = int((got != ref_out[:nv].reshape(nv, -1)).sum()) if args.out_format == "page" else None
                line["verified_sheets"] = nv if bad_bytes == 0 else 0
                line["verify_mismatch_bytes"] = bad_bytes
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
