"""The drop-in proof: the reference's OWN code — backend selector
(imageprocess/backend.c:79-97), L2 wrappers, sheet_process.c, src/core/sheet_stages.c,
lib/perf.c, file.c — compiled with -DUNPAPER_WITH_CUDA=1 and linked against
libunpaper_b200.so (oracle/Makefile target `dropin`), runs process_sheet() with
`--device=cuda` selected, and produces the same sheets as with `--device=cpu`."""
import ctypes as C

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from oracle import checker  # test infrastructure
from util import himg, linesize, noise_image

pytestmark = pytest.mark.gpu
SMALL_BOX = (0.60, 0.72)


@pytest.fixture(scope="module")
def dropin():
    lib = checker.load_dropin()
    if lib is None:
        pytest.skip("oracle/_ref/libunpaper_dropin.so not built")
    yield lib
    lib.ref_select_device(0)


def _both(dropin, cfg, pages, w, h, fmt, threads=1):
    dropin.ref_select_device(1)
    assert dropin.ref_backend_name() == b"cuda"          # the vtable exported by libunpaper_b200.so
    cuda, rc = checker.process_sheets_cpu(dropin, "ref_", cfg, pages, w, h, fmt, threads=threads)
    dropin.ref_select_device(0)
    assert dropin.ref_backend_name() == b"cpu"
    cpu, rp = checker.process_sheets_cpu(dropin, "ref_", cfg, pages, w, h, fmt, threads=8)
    assert all(r.status == 0 for r in rc) and all(r.status == 0 for r in rp)
    diff = cuda != cpu
    assert not diff.any(), f"{int(diff.sum())} differing bytes; per sheet {diff.reshape(len(rc), -1).sum(axis=1)}"
    return cuda


def test_dropin_config2_gray(dropin, ref_lib):
    """BASELINE config 2 (reduced size): default pipeline; also equal to the sheet engine."""
    from unpaper_gpu_b200.lib import Engine
    w, h = 620, 877
    pages = np.stack([synth.gray_page(400 + i, w, h, box=SMALL_BOX) for i in range(3)])
    cfg = U.default_sheet_config()
    out = _both(dropin, cfg, pages, w, h, U.FMT_GRAY8)
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=2, lanes=1)
    eout, _ = eng.process_numpy(pages)
    eng.close()
    assert np.array_equal(out, eout)
    # and equal to the separately built, CUDA-free reference library
    rout, _ = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, U.FMT_GRAY8, threads=8)
    assert np.array_equal(out, rout)


def test_dropin_config3_color(dropin):
    """BASELINE config 3: RGB24, grayfilter + blurfilter + cubic deskew."""
    w, h = 620, 877
    pages = np.stack([synth.color_page(410 + i, w, h) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = 1
    _both(dropin, cfg, pages, w, h, U.FMT_RGB24)


def test_dropin_double_layout_and_options(dropin):
    """--layout double, then a run with wipes, borders, mirror, shift and linear interpolation."""
    w, h = 1754, 1240
    pages = np.stack([synth.double_sheet(420 + i, w, h) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.layout = U.LAYOUT_DOUBLE
    _both(dropin, cfg, pages, w, h, U.FMT_GRAY8)
    w, h = 620, 877
    pages = np.stack([synth.gray_page(430 + i, w, h, box=SMALL_BOX) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.interpolate_type = U.INTERP_LINEAR
    cfg.pre_wipe_count = 1; cfg.pre_wipes[0] = U.rect(100, 120, 160, 170)
    cfg.border = U.Border(10, 0, 0, 12)
    cfg.pre_mirror, cfg.post_shift = U.Direction(True, False), U.Delta(5, -7)
    _both(dropin, cfg, pages, w, h, U.FMT_GRAY8)


def test_dropin_perf_recorder_and_stream_pool(dropin, capfd):
    """lib/perf.c with the CUDA event pairs (cuda_runtime.h:86-88) and the global stream pool
    handed out per job like lib/batch_worker.c:198-255 (4 worker threads)."""
    w, h = 620, 877
    pages = np.stack([synth.gray_page(440 + i, w, h, box=SMALL_BOX) for i in range(6)])
    cfg = U.default_sheet_config()
    dropin.ref_set_perf(1)
    assert dropin.ref_stream_pool(4) == 0
    try:
        _both(dropin, cfg, pages, w, h, U.FMT_GRAY8, threads=4)
        assert dropin.ref_stream_pool_acquisitions() >= 6
    finally:
        dropin.ref_set_perf(0)
        dropin.ref_stream_pool(0)
    txt = capfd.readouterr()
    assert "cuda" in (txt.out + txt.err)        # perf_recorder_print(..., "cuda") ran (sheet_stages.c:689-693)


@pytest.mark.parametrize("fmt", [U.FMT_GRAY8, U.FMT_RGB24])
@pytest.mark.parametrize("mode", [0, 1])
def test_create_image_from_gpu(dropin, ref_ops, fmt, mode):
    """image.h:32-61: a foreign device buffer with an odd pitch becomes an Image, is processed
    in place through the vtable (never reallocated, never overwritten by an upload), read back
    with image_ensure_cpu, and released (owns_memory=true frees it)."""
    w, h = 203, 131
    bpp = 1 if fmt == U.FMT_GRAY8 else 3
    src = np.full((h, linesize(fmt, w)), 255, dtype=np.uint8)
    src[30:100, 50 * bpp:150 * bpp] = noise_image(21, 100, 70, fmt, dark=0.5)[:, :100 * bpp]   # ink inside white margins
    row = U.bytes_per_row(fmt, w)
    pitch = row + 13
    wipe = U.rect(10, 12, 60, 40)
    mp = U.default_sheet_config().mask_detection
    mp.maximum_width, mp.maximum_height = w, h
    mp.scan_size = U.RectangleSize(10, 10)
    mp.minimum_width = mp.minimum_height = 1
    out = np.zeros((h, row), dtype=np.uint8)
    mask = U.Rectangle()
    dropin.ref_select_device(1)
    dropin.ref_gpu_image_roundtrip.argtypes = [C.POINTER(U.HostImage), C.c_int, C.c_int, C.POINTER(U.Rectangle),
                                               C.POINTER(U.MaskDetectionParameters), C.POINTER(U.Rectangle), C.c_void_p]
    n = dropin.ref_gpu_image_roundtrip(C.byref(himg(src, fmt, w)), pitch, mode, C.byref(wipe), C.byref(mp),
                                       C.byref(mask), out.ctypes.data)
    dropin.ref_select_device(0)
    assert n >= 0, "residency protocol violated"
    # the same three ops on the CPU backend
    want = src.copy()
    pts = (U.Point * 1)(U.Point(w // 2, h // 2))
    m = (U.Rectangle * 1)()
    rn = ref_ops.call("detect_masks", C.byref(himg(want, fmt, w)), C.byref(mp), pts, 1, m)
    assert n == rn and U.rect_tuple(mask) == U.rect_tuple(m[0])
    ref_ops.call("wipe_rectangle", C.byref(himg(want, fmt, w)), C.byref(wipe), U.Pixel(10, 20, 30))
    tmp = want.copy()
    ref_ops.call("copy_rectangle", C.byref(himg(tmp, fmt, w)), C.byref(himg(want, fmt, w)), C.byref(wipe), U.Point(110, 72))
    assert np.array_equal(out, want[:, :row])
