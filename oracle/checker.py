"""TEST INFRASTRUCTURE: loaders for the two checker libraries.

  * oracle/_ref/libunpaper_ref.so — the UNMODIFIED reference CPU backend
    (built by oracle/Makefile from /root/reference; `ref_*` entry points);
  * oracle/liboracle.so — this repo's CPU restatement (`orc_*` entry points).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product package never does.
"""
import ctypes as C
import os

import numpy as np

from unpaper_gpu_b200.abi import SheetConfig, SheetResult, bytes_per_row

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ref():
    """oracle/_ref/libunpaper_ref.so (the reference CPU backend) or None."""
    p = os.path.join(ROOT, "oracle", "_ref", "libunpaper_ref.so")
    if not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    lib.ref_process_sheets.argtypes = [
        C.POINTER(SheetConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
        C.POINTER(SheetResult), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_process_sheets.restype = C.c_int
    lib.ref_online_cpus.restype = C.c_int
    return lib


def load_oracle():
    """oracle/liboracle.so (this repo's CPU restatement) or None."""
    p = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    if hasattr(lib, "orc_process_sheets"):
        lib.orc_process_sheets.argtypes = [
            C.POINTER(SheetConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
            C.POINTER(SheetResult), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_process_sheets.restype = C.c_int
    return lib


def process_sheets_cpu(lib, prefix, cfg, pages, page_w, page_h, fmt, threads=1, want_out=True):
    """Run ``n`` sheets through a CPU library's process_sheet() equivalent.

    ``pages``: uint8 array holding n*input_count tightly packed pages.
    Returns (out array [n, sheet_h, sheet_row_bytes] or None, list of SheetResult)."""
    row = bytes_per_row(fmt, page_w)
    per_sheet = row * page_h * cfg.input_count
    pages = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1)
    n = pages.size // per_sheet
    sw, sh = page_w * cfg.input_count, page_h
    out = np.empty((n, sh, bytes_per_row(fmt, sw)), dtype=np.uint8) if want_out else None
    res = (SheetResult * n)()
    w, h = C.c_int(), C.c_int()
    rc = getattr(lib, prefix + "process_sheets")(
        C.byref(cfg), pages.ctypes.data, page_w, page_h, fmt, n,
        out.ctypes.data if want_out else None, res, threads, C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"{prefix}process_sheets: {rc}")
    return out, list(res)
