#!/usr/bin/env python
"""Side measurement: ONE process feeding N GPUs through the library's page scheduler
(unpaper_b200_pool_*: shared job counter, pinned decoded-page ring + producer and feeder
thread per device) — host buffers in, host sheets out, BASELINE config 2 pages.

  python tools/bench_pool.py --gpus 2 [--pages 4096] [--group 32] [--lanes 4]

The producer hook is a C function (compiled here with gcc): run 0 copies real pages into
the pinned slots (and is checked against a single-engine run), the timed runs use a
producer that fills a slot position the first time it sees it and leaves it as it is
afterwards (a decoder that is never the bottleneck).
Time = max over devices of the engine's CUDA-event time for its stream; wall clock beside it."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import unpaper_gpu_b200 as U  # noqa: E402
from unpaper_gpu_b200 import synth  # noqa: E402
from unpaper_gpu_b200 import lib as L  # noqa: E402

HELPER = r"""
#include <stdint.h>
#include <string.h>
typedef struct { const uint8_t *pages; int n; size_t bytes; } Src;
int fill_produce(void *user, int idx, uint8_t *dst) { Src *s = (Src *)user; memcpy(dst, s->pages + (size_t)(idx % s->n) * s->bytes, s->bytes); return 0; }
/* a decoder that is never the bottleneck: every pinned slot position is filled with a real page
 * the first time it is handed out and left as it is afterwards */
#define NSEEN 65536
static uint8_t *seen[NSEEN];
int lazy_produce(void *user, int idx, uint8_t *dst) {
  unsigned h = (unsigned)(((uintptr_t)dst >> 12) * 2654435761u) % NSEEN;
  for (;;) {
    uint8_t *cur = __atomic_load_n(&seen[h], __ATOMIC_ACQUIRE);
    if (cur == dst) return 0;
    if (cur == NULL) {
      uint8_t *expect = NULL;
      if (__atomic_compare_exchange_n(&seen[h], &expect, dst, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) return fill_produce(user, idx, dst);
      continue;
    }
    h = (h + 1) % NSEEN;
  }
}
typedef struct { uint8_t *out; size_t bytes; int keep; long done; } Dst;
int keep_sink(void *user, int idx, int dev, const uint8_t *sheet, const void *res) {
  Dst *d = (Dst *)user; (void)dev; (void)res;
  if (idx < d->keep) memcpy(d->out + (size_t)idx * d->bytes, sheet, d->bytes);
  __sync_fetch_and_add(&d->done, 1);
  return 0;
}
"""


class Src(C.Structure):
    _fields_ = [("pages", C.c_void_p), ("n", C.c_int), ("bytes", C.c_size_t)]


class Dst(C.Structure):
    _fields_ = [("out", C.c_void_p), ("bytes", C.c_size_t), ("keep", C.c_int), ("done", C.c_long)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--pages", type=int, default=4096)
    ap.add_argument("--group", type=int, default=32)
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--distinct", type=int, default=32)
    ap.add_argument("--runs", type=int, default=3)
    a = ap.parse_args()
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "h.c"), "w") as f:
        f.write(HELPER)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", os.path.join(d, "h.so"), os.path.join(d, "h.c")])
    h = C.CDLL(os.path.join(d, "h.so"))
    W, H = synth.A4_W, synth.A4_H
    pages = np.stack([synth.gray_page(i, W, H) for i in range(a.distinct)])
    lib = L.load()
    cfg = U.default_sheet_config()
    pool = L.Pool(cfg, list(range(a.gpus)), W, H, U.FMT_GRAY8, group_pages=a.group, lanes=a.lanes)
    keep = min(a.distinct, a.pages)
    out = np.zeros((keep, H, W), dtype=np.uint8)
    src = Src(pages.ctypes.data, a.distinct, W * H)
    dst = Dst(out.ctypes.data, W * H, keep, 0)
    run = lib.unpaper_b200_pool_run
    run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def go(producer):
        dst.done = 0
        t0 = time.perf_counter()
        rc = run(pool.h, a.pages, C.cast(producer, C.c_void_p), C.byref(src), C.cast(h.keep_sink, C.c_void_p), C.byref(dst), None)
        wall = time.perf_counter() - t0
        assert rc == 0 and dst.done == a.pages, (rc, dst.done, L.last_error())
        lib.unpaper_b200_engine_last_device_ms.restype = C.c_double
        dev_ms = max(lib.unpaper_b200_engine_last_device_ms(C.c_void_p(lib.unpaper_b200_pool_engine(pool.h, i))) for i in range(a.gpus))
        return dev_ms, wall * 1e3, pool.sheets_done()

    go(h.fill_produce)                                   # real pages copied in by the producer; outputs kept
    eng = L.Engine(cfg, W, H, U.FMT_GRAY8, group_pages=8, lanes=2)
    want, _ = eng.process_numpy(pages[:min(keep, 8)])
    eng.close()
    verified = bool(np.array_equal(out[:len(want)], want))
    dst.keep = 0
    best = None
    if os.environ.get("POOL_PROFILE"):
        lib.unpaper_b200_engine_set_profiling(C.c_void_p(lib.unpaper_b200_pool_engine(pool.h, 0)), 1)
    go(h.lazy_produce)                                   # touches every slot position once
    for _ in range(a.runs):
        dev_ms, wall_ms, done = go(h.lazy_produce)
        if best is None or dev_ms < best[0]:
            best = (dev_ms, wall_ms, done)
    print(json.dumps({"metric": "A4 300dpi GRAY8 pages/sec full pipeline, one process, page scheduler", "n_gpus": a.gpus,
                      "pages": a.pages, "group": a.group, "lanes": a.lanes,
                      "pages_per_sec_device_time": round(a.pages / (best[0] / 1e3), 1),
                      "pages_per_sec_wall": round(a.pages / (best[1] / 1e3), 1),
                      "sheets_per_device": best[2], "verified_against_single_engine": verified,
                      "h2d_bytes_per_page": W * H, "d2h_bytes_per_page": W * H}))
    if os.environ.get("POOL_PROFILE"):
        names = (C.c_char_p * 32)(); ms = (C.c_double * 32)(); cnt = (C.c_uint64 * 32)()
        n = lib.unpaper_b200_engine_get_profile(C.c_void_p(lib.unpaper_b200_pool_engine(pool.h, 0)), 32, names, ms, cnt, None)
        print({names[i].decode(): round(ms[i] / max(cnt[i], 1), 3) for i in range(n)}, file=sys.stderr)
    pool.close()


if __name__ == "__main__":
    main()
