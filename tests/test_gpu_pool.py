"""The page scheduler across GPUs (include/unpaper_b200.h layer 4; the role of the
reference's lib/batch_worker.c:174-296 + lib/decode_queue.h) and the engine's stream API.
Single process; uses GPUs 0 and 1 when the box has two, else two engines on GPU 0 —
the scheduling logic (shared job counter, pinned slot ring, producer / feeder threads)
is the same."""
import ctypes as C

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from oracle import checker  # test infrastructure

pytestmark = pytest.mark.gpu
SMALL_BOX = (0.60, 0.72)


def _devices():
    from unpaper_gpu_b200 import lib as L
    return [0, 1] if L.load().unpaper_b200_device_count() >= 2 else [0, 0]


def test_pool_two_devices_vs_reference(ref_lib):
    from unpaper_gpu_b200.lib import Pool
    w, h, n = 620, 877, 23
    pages = np.stack([synth.gray_page(500 + i, w, h, box=SMALL_BOX) for i in range(n)])
    cfg = U.default_sheet_config()
    cfg.no_deskew_sheets = U.multi_index([5, 17])            # per-sheet switches follow the JOB index, not the device-local order
    pool = Pool(cfg, _devices(), w, h, U.FMT_GRAY8, group_pages=3, lanes=2)
    out = np.zeros((n, h, w), dtype=np.uint8)
    seen, produced = [], []

    def produce(idx, dst):
        produced.append(idx)
        C.memmove(dst, pages[idx].ctypes.data, w * h)
        return 0

    def sink(idx, dev, ptr, res):
        seen.append((idx, dev, res.status, res.deskew_mask_count))
        C.memmove(out[idx].ctypes.data, ptr, pool.sheet_bytes)
        return 0

    res = (U.SheetResult * n)()
    pool.run(n, produce, sink, res)
    done = pool.sheets_done()
    assert sorted(produced) == list(range(n)) and sorted(s[0] for s in seen) == list(range(n))
    assert sum(done) == n and all(d > 0 for d in done), done          # both engines took part
    # second run on the same pool (slots and engines are reused)
    out2 = np.zeros_like(out)
    pool.run(n, produce, lambda i, d, p, r: C.memmove(out2[i].ctypes.data, p, pool.sheet_bytes) and 0)
    pool.close()
    rout, rres = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, U.FMT_GRAY8, threads=8)
    assert np.array_equal(out, rout) and np.array_equal(out2, rout)
    for i in range(n):
        assert res[i].status == 0 and res[i].rotation[0] == rres[i].rotation[0]
        assert res[i].deskew_mask_count == rres[i].deskew_mask_count == (0 if i + 1 in (5, 17) else 1)


def test_pool_producer_end_and_failure():
    """The producer ends the job list early (return 1) or fails (< 0)."""
    from unpaper_gpu_b200.lib import Pool
    w, h = 620, 877
    pages = np.stack([synth.gray_page(530 + i, w, h, box=SMALL_BOX) for i in range(4)])
    pool = Pool(U.default_sheet_config(), _devices(), w, h, U.FMT_GRAY8, group_pages=2, lanes=1)
    got = []

    def produce(idx, dst):
        if idx >= 7:
            return 1
        C.memmove(dst, pages[idx % 4].ctypes.data, w * h)
        return 0

    pool.run(100, produce, lambda i, d, p, r: got.append(i) or 0)
    assert sorted(got) == list(range(7))
    with pytest.raises(RuntimeError, match="-4"):
        pool.run(10, lambda idx, dst: -1 if idx == 3 else produce(idx, dst), lambda i, d, p, r: 0)
    got.clear()
    pool.run(5, produce, lambda i, d, p, r: got.append(i) or 0)      # still usable
    assert sorted(got) == list(range(5))
    pool.close()


def test_engine_stream_feed_matches_process():
    """begin / feed / feed / end gives the same bytes as one process_host call."""
    from unpaper_gpu_b200.lib import Engine
    w, h, n = 620, 877, 11
    pages = np.stack([synth.gray_page(540 + i, w, h, box=SMALL_BOX) for i in range(n)])
    eng = Engine(U.default_sheet_config(), w, h, U.FMT_GRAY8, group_pages=2, lanes=2)
    want, wres = eng.process_numpy(pages)
    out = np.zeros_like(want)
    res = (U.SheetResult * n)()
    order = []
    eng.set_sheet_callback(lambda idx, ptr, r: order.append(idx) or 0)
    eng.stream_begin(True)
    cuts = [0, 1, 6, 6, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        eng.stream_feed(pages[a:].ctypes.data, out[a:].ctypes.data, b - a, C.cast(C.byref(res, a * C.sizeof(U.SheetResult)), C.POINTER(U.SheetResult)))
    eng.stream_end()
    eng.set_sheet_callback(None)
    eng.close()
    assert order == list(range(n))
    assert np.array_equal(out, want)
    assert [r.rotation[0] for r in res] == [r.rotation[0] for r in wres]
