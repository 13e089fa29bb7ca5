#!/usr/bin/env python
"""Side measurement (not the headline): BASELINE configs 3 and 4 through the engine,
sheets resident in HBM.  One JSON line per config."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaper_gpu_b200 as U  # noqa: E402
from unpaper_gpu_b200 import synth  # noqa: E402
from unpaper_gpu_b200.lib import Engine  # noqa: E402

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def run(name, cfg, w, h, fmt, pages, n, group, lanes):
    eng = Engine(cfg, w, h, fmt, group_pages=group, lanes=lanes)
    distinct = len(pages)
    dev_in = torch.from_numpy(np.stack([pages[i % distinct] for i in range(n)])).cuda()
    dev_out = torch.empty((n, eng.sheet_bytes), dtype=torch.uint8, device="cuda")
    res = (U.SheetResult * n)()
    for _ in range(2):
        eng.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), n, False, res)
    ms = 0.0
    for _ in range(3):
        eng.process_ptr(dev_in.data_ptr(), dev_out.data_ptr(), n, False, res)
        ms += eng.last_device_ms()
    bad = sum(1 for r in res if r.status != 0)
    eng.close()
    print(json.dumps({"config": name, "sheets_per_sec": round(3 * n / (ms / 1e3), 1), "sheets_per_step": n,
                      "sheet": f"{w}x{h}", "failed_sheets": bad}))


cfg3 = U.default_sheet_config()
cfg3.no_blackfilter = cfg3.no_noisefilter = 1
run("C3 colour A4 RGB24: grayfilter + blurfilter + cubic deskew", cfg3, synth.A4_W, synth.A4_H, U.FMT_RGB24,
    [synth.color_page(i) for i in range(4)], 256, 16, 4)
cfg4 = U.default_sheet_config()
cfg4.layout = U.LAYOUT_DOUBLE
run("C4 double-600 GRAY8 7016x4960, layout double", cfg4, 7016, 4960, U.FMT_GRAY8,
    [synth.double_sheet(i) for i in range(4)], 128, 8, 4)
