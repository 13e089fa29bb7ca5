#!/usr/bin/env python
"""Side measurement (not the headline): the same A4 job fed as 1-bit pages (pbm scans, the
reference's PDF path) — H2D and D2H carry 1 bit/px, the device expands and re-packs.
Prints one JSON line; pages/s through host buffers and HBM-resident."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unpaper_gpu_b200 as U  # noqa: E402
from unpaper_gpu_b200 import synth  # noqa: E402
from unpaper_gpu_b200.lib import Engine  # noqa: E402

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
N, W, H = int(os.environ.get("BENCH_PAGES", "1024")), synth.A4_W, synth.A4_H
distinct = [np.packbits(synth.gray_page(i, W, H) < 128, axis=1) for i in range(8)]
host = torch.from_numpy(np.stack([distinct[i % 8] for i in range(N)])).pin_memory()
eng = Engine(U.default_sheet_config(), W, H, U.FMT_MONOWHITE, group_pages=32, lanes=8)
out_host = torch.empty((N, eng.sheet_bytes), dtype=torch.uint8).pin_memory()
dev_in, dev_out = host.cuda(), torch.empty((N, eng.sheet_bytes), dtype=torch.uint8, device="cuda")
res = (U.SheetResult * N)()
r = {}
for name, a, b, hostmode in (("hbm_resident", dev_in, dev_out, False), ("e2e", host, out_host, True)):
    for _ in range(2):
        eng.process_ptr(a.data_ptr(), b.data_ptr(), N, hostmode, res)
    ms = 0.0
    for _ in range(3):
        eng.process_ptr(a.data_ptr(), b.data_ptr(), N, hostmode, res)
        ms += eng.last_device_ms()
    r[name] = round(3 * N / (ms / 1e3), 1)
bad = sum(1 for x in res if x.status != 0)
print(json.dumps({"metric": "pages_per_sec_a4_300dpi_1bit", "pages_per_step": N, "hbm_resident": r["hbm_resident"],
                  "e2e": r["e2e"], "h2d_bytes_per_page": W // 8 * H, "d2h_bytes_per_page": eng.sheet_bytes,
                  "failed_sheets": bad}))
