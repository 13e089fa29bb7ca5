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


def process_sheets_cpu(lib, prefix, cfg, pages, page_w, page_h, fmt, threads=1, want_out=True, out_size=None):
    """Run ``n`` sheets through a CPU library's process_sheet() equivalent.

    ``pages``: uint8 array holding n*input_count tightly packed pages.  ``out_size``: (w, h) of
    the finished sheets when size-changing options are set (default: pages side by side).
    Returns (out array [n, sheet_h, sheet_row_bytes] or None, list of SheetResult)."""
    row = bytes_per_row(fmt, page_w)
    per_sheet = row * page_h * cfg.input_count
    pages = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1)
    n = pages.size // per_sheet
    sw, sh = out_size if out_size else (page_w * cfg.input_count, page_h)
    out = np.empty((n, sh, bytes_per_row(fmt, sw)), dtype=np.uint8) if want_out else None
    res = (SheetResult * n)()
    w, h = C.c_int(sw), C.c_int(sh)
    rc = getattr(lib, prefix + "process_sheets")(
        C.byref(cfg), pages.ctypes.data, page_w, page_h, fmt, n,
        out.ctypes.data if want_out else None, res, threads, C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"{prefix}process_sheets: {rc}")
    return out, list(res)


def read_pnm(path):
    """(format, width, height, uint8 array [h, row_bytes]) of a P4 / P5 / P6 file as the
    reference's saveImageDirect() writes it (file.c:134-176)."""
    from unpaper_gpu_b200.abi import FMT_GRAY8, FMT_MONOWHITE, FMT_RGB24
    with open(path, "rb") as f:
        data = f.read()
    magic = data[:2]
    fmt = {b"P4": FMT_MONOWHITE, b"P5": FMT_GRAY8, b"P6": FMT_RGB24}[magic]
    fields, pos = [], 2
    need = 2 if magic == b"P4" else 3
    while len(fields) < need:
        while data[pos:pos + 1].isspace():
            pos += 1
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        fields.append(int(data[pos:end]))
        pos = end
    pos += 1   # the single whitespace byte after the header
    w, h = fields[0], fields[1]
    row = bytes_per_row(fmt, w)
    return fmt, w, h, np.frombuffer(data, dtype=np.uint8, count=row * h, offset=pos).reshape(h, row).copy()


def process_sheets_files_cpu(lib, cfg, pages, page_w, page_h, fmt, out_fmt=-1, output_count=1, threads=1):
    """The reference's process_sheet() WITH its output stage (sheet_stages.c:536-631 ->
    saveImage, file.c:186-262).  Returns ([per sheet: [per output page: (fmt, w, h, array)]], results)."""
    import shutil
    import tempfile
    row = bytes_per_row(fmt, page_w)
    per_sheet = row * page_h * cfg.input_count
    pages = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1)
    n = pages.size // per_sheet
    res = (SheetResult * n)()
    lib.ref_process_sheets_files.argtypes = [
        C.POINTER(SheetConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_char_p, C.POINTER(SheetResult), C.c_int]
    lib.ref_process_sheets_files.restype = C.c_int
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="unpaper_ref_", dir=base)
    try:
        rc = lib.ref_process_sheets_files(C.byref(cfg), pages.ctypes.data, page_w, page_h, fmt, n, out_fmt,
                                          output_count, d.encode(), res, threads)
        if rc != 0:
            raise RuntimeError(f"ref_process_sheets_files: {rc}")
        out = [[read_pnm(os.path.join(d, f"s{i:06d}_{j}.pnm")) for j in range(output_count)] for i in range(n)]
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out, list(res)


def save_image_cpu(lib, himg, out_fmt):
    """The reference's saveImage() (file.c:186-262) on one image; returns (fmt, w, h, array)."""
    import tempfile
    from unpaper_gpu_b200.abi import HostImage
    lib.ref_save_image.argtypes = [C.POINTER(HostImage), C.c_int, C.c_char_p]
    lib.ref_save_image.restype = C.c_int
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    fd, path = tempfile.mkstemp(prefix="unpaper_ref_", suffix=".pnm", dir=base)
    os.close(fd)
    try:
        lib.ref_save_image(C.byref(himg), out_fmt, path.encode())
        return read_pnm(path)
    finally:
        os.unlink(path)


def load_dropin():
    """oracle/_ref/libunpaper_dropin.so: the reference's own L1-L3 built with
    -DUNPAPER_WITH_CUDA=1 and LINKED AGAINST libunpaper_b200.so (oracle/Makefile target
    `dropin`).  Same `ref_*` entry points as load_ref() plus ref_select_device()."""
    p = os.path.join(ROOT, "oracle", "_ref", "libunpaper_dropin.so")
    if not os.path.exists(p):
        return None
    lib = C.CDLL(p)
    lib.ref_process_sheets.argtypes = [
        C.POINTER(SheetConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
        C.POINTER(SheetResult), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_process_sheets.restype = C.c_int
    lib.ref_backend_name.restype = C.c_char_p
    return lib
