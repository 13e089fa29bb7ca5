"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares;
struct layouts agree between C and the ctypes mirror; no compute calls."""
import ctypes as C
import os
import re
import subprocess

import unpaper_gpu_b200 as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "unpaper_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:unpaper_b200_|unpaper_cuda_|cuda_stream_pool_|image_|create_image_from_gpu)\w*)\s*\(", text))
    names.add("backend_cuda")
    for pool in ("", "integral_", "scratch_"):      # B200_DECL_GLOBAL_POOL(...)
        for fn in ("init", "cleanup", "active", "acquire", "release", "get_stats", "print_stats"):
            names.add(f"cuda_mempool_{pool}global_{fn}")
    names |= {"cuda_mempool_create", "cuda_mempool_destroy", "cuda_mempool_acquire", "cuda_mempool_release"}
    return sorted(n for n in names if not n.endswith("##"))


def test_library_exports_every_declared_symbol():
    from unpaper_gpu_b200 import lib as L
    handle = L.load()
    missing = [n for n in _declared_symbols() if not hasattr(handle, n)]
    assert not missing, missing
    assert len(_declared_symbols()) > 50


def test_library_exports_nothing_else():
    """-fvisibility via the linker map (csrc/exports.map): only the declared API and the
    reference's boundary names are dynamic symbols — no stage_*, no kernel launchers, no cudart."""
    from unpaper_gpu_b200 import lib as L
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    syms = [ln.split()[-1] for ln in out.splitlines() if ln.strip()]
    allowed = re.compile(r"^(backend_cuda|image_\w+|create_image_from_gpu|unpaper_cuda_\w+|unpaper_b200_\w+|"
                         r"cuda_stream_pool_\w+|cuda_mempool_\w+)$")
    extra = [s for s in syms if not allowed.match(s)]
    assert not extra, extra[:20]
    assert len(syms) < 160


def test_backend_vtable_shape():
    """`backend_cuda` = name + 20 function pointers (reference backend.h:19-57)."""
    from unpaper_gpu_b200 import lib as L
    handle = L.load()
    tab = (C.c_void_p * 21).in_dll(handle, "backend_cuda")
    assert C.cast(tab[0], C.c_char_p).value == b"cuda"
    assert all(tab[i] for i in range(1, 21))


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "unpaper_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(B200SheetConfig),sizeof(B200SheetResult),sizeof(B200HostImage),sizeof(BlackfilterParameters),'
                   'sizeof(MaskDetectionParameters),offsetof(B200SheetConfig,deskew),offsetof(B200SheetResult,borders));}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(U.SheetConfig), C.sizeof(U.SheetResult), C.sizeof(U.HostImage), C.sizeof(U.BlackfilterParameters),
            C.sizeof(U.MaskDetectionParameters), U.SheetConfig.deskew.offset, U.SheetResult.borders.offset]
    assert got == want


def test_defaults_match_c():
    from unpaper_gpu_b200 import lib as L
    c = U.SheetConfig()
    L.load().unpaper_b200_sheet_config_defaults(C.byref(c))
    p = U.default_sheet_config()
    assert bytes(c) == bytes(p)


def test_product_does_not_touch_oracle():
    """The product tree must not reference anything under oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "unpaper-gpu_b200")):
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py")) or f == "Makefile":
                if "oracle" in open(os.path.join(dp, f), errors="ignore").read().lower().replace("oracle/makefile", ""):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_output_side_host_functions(tmp_path):
    """The parts of the output side that are plain host code: saveImage()'s format
    mapping (file.c:201-208) and saveImageDirect's PNM layout (file.c:134-176)."""
    import numpy as np
    from unpaper_gpu_b200 import lib as L
    lib = L.load()
    assert lib.unpaper_b200_output_format(U.FMT_Y400A) == U.FMT_GRAY8
    assert lib.unpaper_b200_output_format(U.FMT_MONOBLACK) == U.FMT_MONOWHITE
    assert lib.unpaper_b200_output_format(U.FMT_GRAY8) == U.FMT_GRAY8
    buf = C.create_string_buffer(64)
    for fmt, want in ((U.FMT_GRAY8, b"P5\n31 7\n255\n"), (U.FMT_RGB24, b"P6\n31 7\n255\n"), (U.FMT_MONOWHITE, b"P4\n31 7\n")):
        n = lib.unpaper_b200_pnm_header(fmt, 31, 7, buf, 64)
        assert buf.raw[:n] == want
    assert lib.unpaper_b200_pnm_header(U.FMT_Y400A, 31, 7, buf, 64) < 0
    assert lib.unpaper_b200_pnm_header(U.FMT_GRAY8, 31, 7, buf, 4) < 0          # buffer too small
    # rows with padding are written tight
    img = np.arange(5 * 16, dtype=np.uint8).reshape(5, 16)
    path = os.path.join(str(tmp_path), "t.pgm")
    assert lib.unpaper_b200_write_pnm(path.encode(), img.ctypes.data, 16, 13, 5, U.FMT_GRAY8) == 0
    blob = open(path, "rb").read()
    assert blob == b"P5\n13 5\n255\n" + img[:, :13].tobytes()
    assert lib.unpaper_b200_write_pnm(b"/nonexistent-dir/x.pgm", img.ctypes.data, 16, 13, 5, U.FMT_GRAY8) < 0


def test_library_carries_sm_100a_code_with_line_info():
    """The in-tree library is built for sm_100a (SASS + compute_100a PTX) with -lineinfo and
    without FMA contraction, as csrc/Makefile says."""
    import shutil
    mk = open(os.path.join(ROOT, "unpaper-gpu_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=[sm_100a,compute_100a]" in mk and "-lineinfo" in mk and "--fmad=false" in mk
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "unpaper-gpu_b200", "libunpaper_b200.so")
    elfs = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    kernels = [l for l in elfs.splitlines() if "sm_100a" in l]
    assert len(kernels) >= 6, elfs          # one cubin per k_*.cu


def test_swar_byte_compare(tmp_path):
    """csrc/swar.h (the byte compares of the row / rectangle / cell statistics kernels) compiled as C:
    every (lo, hi) pair over every byte value, four different bytes per word, against the plain
    comparison — and no bit other than bit 7 of a byte is ever set."""
    src = tmp_path / "swar_check.c"
    src.write_text(r'''
#include <stdio.h>
#include "swar.h"
int main(void) {
  long bad = 0;
  for (int lo = -3; lo <= 300; lo++)
    for (int hi = -3; hi <= 300; hi++) {
      Lt4 a = lt4_make(lo), b = lt4_make(hi + 1);
      for (int v = 0; v < 256; v++) {
        unsigned w = (unsigned)v | ((unsigned)(255 - v) << 8) | ((unsigned)((v * 7 + 3) & 255) << 16) | ((unsigned)((v * 13 + 101) & 255) << 24);
        unsigned m = range4_bit7(w, a, b), l = lt4(w, b);
        for (int k = 0; k < 4; k++) {
          int byte = (w >> (8 * k)) & 255;
          if (((m >> (8 * k + 7)) & 1) != (byte >= lo && byte <= hi)) bad++;
          if (((l >> (8 * k + 7)) & 1) != (byte <= hi)) bad++;
          if (((m | l) >> (8 * k)) & 0x7F) bad++;
        }
      }
    }
  printf("%ld\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "swar_check"
    csrc = os.path.join(os.path.dirname(U.__file__), "csrc")
    if not os.path.isdir(csrc):
        csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unpaper-gpu_b200", "csrc")
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-I", csrc, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout
