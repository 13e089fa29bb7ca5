"""Loader and thin wrappers for libunpaper_b200.so (the CUDA product).

There is no CPU fallback: if the shared library is missing, or no GPU is
present when an op is called, this fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import (HostOps, SheetConfig, SheetResult, bytes_per_row)

# B200SheetDoneFn (include/unpaper_b200.h): user, sheet index, sheet bytes, result
SHEET_DONE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(SheetResult))
# B200PageProducerFn: user, sheet index, pinned destination; B200PoolSheetFn: user, sheet index, device, sheet, result
PAGE_PRODUCER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)
POOL_SHEET_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(SheetResult))

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libunpaper_b200.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the B200 backend has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.unpaper_b200_last_error.restype = C.c_char_p
    lib.unpaper_b200_version.restype = C.c_char_p
    lib.unpaper_cuda_try_init.restype = C.c_int
    lib.unpaper_b200_device_count.restype = C.c_int
    lib.unpaper_b200_set_device.argtypes = [C.c_int]
    lib.unpaper_b200_sheet_config_defaults.argtypes = [C.POINTER(SheetConfig)]
    lib.unpaper_b200_engine_create.argtypes = [C.POINTER(SheetConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.unpaper_b200_engine_create.restype = C.c_void_p
    lib.unpaper_b200_engine_destroy.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_sheet_width.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_sheet_height.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_sheet_bytes.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_sheet_bytes.restype = C.c_size_t
    lib.unpaper_b200_engine_output_width.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_output_count.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_set_first_sheet_nr.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_engine_set_first_sheet_nr.restype = None
    for name in ("unpaper_b200_engine_process_device", "unpaper_b200_engine_process_host"):
        f = getattr(lib, name)
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SheetResult)]
        f.restype = C.c_int
    lib.unpaper_b200_engine_set_output_format.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_engine_output_format.argtypes = [C.c_void_p]
    lib.unpaper_b200_output_format.argtypes = [C.c_int]
    lib.unpaper_b200_pnm_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    lib.unpaper_b200_write_pnm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.unpaper_b200_engine_set_sheet_callback.argtypes = [C.c_void_p, SHEET_DONE_FN, C.c_void_p]
    lib.unpaper_b200_engine_set_sheet_callback.restype = None
    lib.unpaper_b200_engine_launch_count.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_launch_count.restype = C.c_uint64
    lib.unpaper_b200_engine_last_device_ms.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_last_device_ms.restype = C.c_double
    lib.unpaper_b200_engine_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_engine_get_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                                    C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.unpaper_b200_engine_get_profile_spread.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.unpaper_b200_engine_stream_begin.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_engine_stream_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SheetResult), C.c_int]
    lib.unpaper_b200_engine_stream_poll.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_stream_in_flight.argtypes = [C.c_void_p]
    lib.unpaper_b200_engine_stream_end.argtypes = [C.c_void_p]
    lib.unpaper_b200_pool_create.argtypes = [C.POINTER(SheetConfig), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int]
    lib.unpaper_b200_pool_create.restype = C.c_void_p
    lib.unpaper_b200_pool_destroy.argtypes = [C.c_void_p]
    lib.unpaper_b200_pool_run.argtypes = [C.c_void_p, C.c_int, PAGE_PRODUCER_FN, C.c_void_p, POOL_SHEET_FN, C.c_void_p,
                                          C.POINTER(SheetResult)]
    lib.unpaper_b200_pool_sheet_bytes.argtypes = [C.c_void_p]
    lib.unpaper_b200_pool_sheet_bytes.restype = C.c_size_t
    lib.unpaper_b200_pool_sheets_done.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_pool_sheets_done.restype = C.c_uint64
    lib.unpaper_b200_pool_engine.argtypes = [C.c_void_p, C.c_int]
    lib.unpaper_b200_pool_engine.restype = C.c_void_p
    _lib = lib
    return lib


def host_ops():
    """The 21 `unpaper_b200_host_*` entry points (CUDA)."""
    return HostOps(load(), "unpaper_b200_host_")


def last_error():
    return load().unpaper_b200_last_error().decode()


class Engine:
    """Python face of the sheet engine (include/unpaper_b200.h layer 3)."""

    def __init__(self, cfg, page_w, page_h, fmt, group_pages=32, lanes=2, device=0):
        self.lib = load()
        self.cfg = cfg
        self.h = self.lib.unpaper_b200_engine_create(C.byref(cfg), device, page_w, page_h, fmt, group_pages, lanes)
        if not self.h:
            raise RuntimeError("engine_create failed: " + last_error())
        self.page_w, self.page_h, self.fmt = page_w, page_h, fmt
        self.sheet_w = self.lib.unpaper_b200_engine_sheet_width(self.h)
        self.sheet_h = self.lib.unpaper_b200_engine_sheet_height(self.h)
        self.sheet_bytes = self.lib.unpaper_b200_engine_sheet_bytes(self.h)
        self.out_count = self.lib.unpaper_b200_engine_output_count(self.h)
        self.out_w = self.lib.unpaper_b200_engine_output_width(self.h)
        self.page_bytes = bytes_per_row(fmt, page_w) * page_h
        self.sheet_in_bytes = self.page_bytes * cfg.input_count
        self.out_fmt = self.lib.unpaper_b200_engine_output_format(self.h)

    def close(self):
        if self.h:
            self.lib.unpaper_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_ptr(self, in_ptr, out_ptr, n, host, results=None):
        fn = self.lib.unpaper_b200_engine_process_host if host else self.lib.unpaper_b200_engine_process_device
        rc = fn(self.h, in_ptr, out_ptr, n, results)
        if rc != 0:
            raise RuntimeError(f"engine process failed ({rc}): {last_error()}")

    def process_numpy(self, pages):
        """pages: uint8 array with n*input_count tightly packed pages (host).
        Returns (out [n, sheet_h, row_bytes] — [n, output_count, sheet_h, row_bytes] for a
        split sheet —, [SheetResult])."""
        pages = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1)
        n = pages.size // self.sheet_in_bytes
        row = self.sheet_bytes // self.sheet_h // self.out_count
        out = np.empty((n, self.sheet_h, row) if self.out_count == 1 else (n, self.out_count, self.sheet_h, row), dtype=np.uint8)
        res = (SheetResult * n)()
        self.process_ptr(pages.ctypes.data, out.ctypes.data, n, True, res)
        return out, list(res)

    def set_output_format(self, fmt):
        """Device-side sheet_stage_output conversion (e.g. FMT_MONOWHITE for .pbm); -1 = page format."""
        if self.lib.unpaper_b200_engine_set_output_format(self.h, fmt) != 0:
            raise RuntimeError("set_output_format failed: " + last_error())
        self.sheet_bytes = self.lib.unpaper_b200_engine_sheet_bytes(self.h)
        self.out_fmt = self.lib.unpaper_b200_engine_output_format(self.h)

    def set_sheet_callback(self, fn):
        """fn(sheet_index, sheet_ptr, result) -> int, called as each sheet completes
        (the reference's post_process_fn); None removes it."""
        if fn is None:
            self._cb = SHEET_DONE_FN(0)
        else:
            self._cb = SHEET_DONE_FN(lambda user, idx, ptr, res: int(fn(idx, ptr, res.contents) or 0))
        self.lib.unpaper_b200_engine_set_sheet_callback(self.h, self._cb, None)

    def launch_count(self):
        return int(self.lib.unpaper_b200_engine_launch_count(self.h))

    def last_device_ms(self):
        return float(self.lib.unpaper_b200_engine_last_device_ms(self.h))

    def set_profiling(self, on):
        self.lib.unpaper_b200_engine_set_profiling(self.h, 1 if on else 0)

    def profile(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        cnt = (C.c_uint64 * 32)()
        n = self.lib.unpaper_b200_engine_get_profile(self.h, 32, names, ms, cnt, None)
        return {names[i].decode(): (ms[i], int(cnt[i])) for i in range(n)}

    def profile_spread(self):
        """stage -> (fastest group ms, slowest group ms) since profiling was switched on."""
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        cnt = (C.c_uint64 * 32)()
        lo, hi = (C.c_double * 32)(), (C.c_double * 32)()
        n = self.lib.unpaper_b200_engine_get_profile(self.h, 32, names, ms, cnt, None)
        self.lib.unpaper_b200_engine_get_profile_spread(self.h, 32, lo, hi)
        return {names[i].decode(): (lo[i], hi[i]) for i in range(n) if cnt[i]}

    # streams: begin, feed batches as they arrive, end (include/unpaper_b200.h)
    def stream_begin(self, host=True):
        if self.lib.unpaper_b200_engine_stream_begin(self.h, 1 if host else 0) != 0:
            raise RuntimeError("stream_begin failed: " + last_error())

    def stream_feed(self, in_ptr, out_ptr, n, results=None, first_index=-1):
        if self.lib.unpaper_b200_engine_stream_feed(self.h, in_ptr, out_ptr, n, results, first_index) != 0:
            raise RuntimeError("stream_feed failed: " + last_error())

    def stream_end(self):
        rc = self.lib.unpaper_b200_engine_stream_end(self.h)
        if rc != 0:
            raise RuntimeError(f"stream_end failed ({rc}): {last_error()}")


class Pool:
    """Python face of the page scheduler across GPUs (include/unpaper_b200.h layer 4)."""

    def __init__(self, cfg, devices, page_w, page_h, fmt, group_pages=32, lanes=2, slot_sheets=0, slots=0):
        self.lib = load()
        devs = (C.c_int * len(devices))(*devices)
        self.h = self.lib.unpaper_b200_pool_create(C.byref(cfg), devs, len(devices), page_w, page_h, fmt, group_pages, lanes,
                                                   slot_sheets, slots)
        if not self.h:
            raise RuntimeError("pool_create failed: " + last_error())
        self.devices = list(devices)
        self.sheet_bytes = self.lib.unpaper_b200_pool_sheet_bytes(self.h)
        self.sheet_in_bytes = bytes_per_row(fmt, page_w) * page_h * cfg.input_count

    def close(self):
        if self.h:
            self.lib.unpaper_b200_pool_destroy(self.h)
            self.h = None

    def run(self, n_sheets, produce, sink, results=None):
        """produce(sheet_index, dst_ptr) -> 0 / 1 (end) / <0; sink(sheet_index, device, sheet_ptr, result) -> int."""
        pcb = PAGE_PRODUCER_FN(lambda user, idx, dst: int(produce(idx, dst) or 0))
        scb = POOL_SHEET_FN(lambda user, idx, dev, ptr, res: int(sink(idx, dev, ptr, res.contents) or 0))
        rc = self.lib.unpaper_b200_pool_run(self.h, n_sheets, pcb, None, scb, None, results)
        if rc != 0:
            raise RuntimeError(f"pool_run failed ({rc}): {last_error()}")

    def sheets_done(self):
        return [int(self.lib.unpaper_b200_pool_sheets_done(self.h, i)) for i in range(len(self.devices))]
