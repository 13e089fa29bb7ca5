#!/usr/bin/env python
"""What each stage costs INSIDE the throughput arm (all lanes overlapping): the default pipeline
is run over HBM-resident pages with one stage switched off at a time; the difference of the
per-page time to the full pipeline is that stage's marginal cost under overlap (a stage whose
kernels are latency-bound and hide behind the other lanes costs less than its isolated time).

usage (GPU box): python tools/stage_ablation.py [--pages 1024] [--group 32] [--lanes 8] > gpurun_out/ablation.json"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (page generator, geometry)
import unpaper_gpu_b200 as U  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=1024)
    ap.add_argument("--group", type=int, default=32)
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--distinct", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    distinct = bench.make_pages(args.distinct, 0)
    import torch
    from unpaper_gpu_b200.lib import Engine
    W, H = bench.W, bench.H
    reps = (args.pages + args.distinct - 1) // args.distinct
    dev_in = torch.from_numpy(distinct).cuda().repeat((reps, 1, 1))[:args.pages].contiguous()
    dev_out = torch.empty((args.pages, H, W), dtype=torch.uint8, device="cuda")
    switches = [None, "no_blackfilter", "no_noisefilter", "no_blurfilter", "no_grayfilter", "no_deskew",
                "no_mask_center", "no_border_scan", "no_border_align"]
    out = {}
    base = None
    for sw in switches:
        cfg = U.default_sheet_config()
        if sw:
            setattr(cfg, sw, 1)
        eng = Engine(cfg, W, H, U.FMT_GRAY8, group_pages=args.group, lanes=args.lanes, device=0)
        for _ in range(2):
            eng.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), args.pages, False, None)
        ms = 0.0
        for _ in range(args.steps):
            eng.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), args.pages, False, None)
            ms += eng.last_device_ms()
        eng.close()
        us = 1000.0 * ms / (args.steps * args.pages)
        if sw is None:
            base = us
        out[sw or "full"] = {"us_per_page": round(us, 2), "pages_per_s": round(1e6 / us, 1),
                             "marginal_us": None if sw is None else round(base - us, 2)}
    print(json.dumps({"pages": args.pages, "group": args.group, "lanes": args.lanes, "ablation": out}, indent=1))


if __name__ == "__main__":
    main()
