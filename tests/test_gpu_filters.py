"""CUDA vs reference CPU backend, filters and detectors — bit-exact.
Mirrors reference tests/cuda_filters_test.c (noise/black/gray/blur on tiny
images, :92-391) and adds what the reference never pins: page-scale inputs."""
import ctypes as C

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from util import (FMTS_ALL, FMTS_BYTE, assert_same, blobs_image, from_gray, himg, noise_image, run_inplace)

pytestmark = pytest.mark.gpu


def _noisy(seed, w, h, fmt, density, edge=False):
    img = noise_image(seed, w, h, fmt, dark=density, lo=0, hi=256)
    if edge and fmt == U.FMT_GRAY8:
        rng = np.random.Generator(np.random.PCG64(seed + 99))
        m = rng.random((h, 12)) < 0.3
        img[:, :12][m] = 0
        m = rng.random((12, w)) < 0.3
        img[:12, :w][m] = 0
    return img


@pytest.mark.parametrize("fmt", FMTS_ALL)
@pytest.mark.parametrize("intensity", [1, 2, 4, 6, 9])
def test_noisefilter_random(cuda_ops, ref_ops, fmt, intensity):
    for seed, (w, h), dens in ((1, (64, 48), 0.02), (2, (241, 179), 0.01), (3, (300, 200), 0.06), (4, (257, 131), 0.15)):
        img = _noisy(seed, w, h, fmt, dens, edge=True)
        a = run_inplace(cuda_ops, "noisefilter", img, fmt, w, intensity, 229)
        b = run_inplace(ref_ops, "noisefilter", img, fmt, w, intensity, 229)
        assert_same(a, b, fmt, w, f"noisefilter I={intensity} seed={seed}")


def test_noisefilter_kats(cuda_ops, ref_ops):
    """Hand-made cases of reference tests/cuda_filters_test.c:92-247."""
    g = np.full((16, 16), 255, dtype=np.uint8)
    g[3, 3] = 0                      # isolated pixel: removed
    g[8:11, 8:11] = 0                # 3x3 block: kept
    g[12, 2] = 0; g[13, 3] = 0       # diagonal pair
    for fmt in (U.FMT_GRAY8, U.FMT_RGB24, U.FMT_Y400A):
        img = from_gray(g, fmt)
        a = run_inplace(cuda_ops, "noisefilter", img, fmt, 16, 4, 229)
        b = run_inplace(ref_ops, "noisefilter", img, fmt, 16, 4, 229)
        assert_same(a, b, fmt, 16, "noisefilter KAT")
        a2 = run_inplace(cuda_ops, "noisefilter", img, fmt, 16, 4, 229)
        assert_same(a, a2, fmt, 16, "noisefilter determinism")


@pytest.mark.parametrize("fmt", [U.FMT_GRAY8, U.FMT_RGB24])
def test_noisefilter_page(cuda_ops, ref_ops, fmt):
    g = synth.gray_page(3, 1240, 1754)
    img = from_gray(g, fmt)
    a = run_inplace(cuda_ops, "noisefilter", img, fmt, 1240, 4, 229)
    b = run_inplace(ref_ops, "noisefilter", img, fmt, 1240, 4, 229)
    assert_same(a, b, fmt, 1240, "noisefilter page")


def _bf_params(w, h, excl=None, intensity=20, depth=500):
    p = U.BlackfilterParameters()
    p.scan_size = U.RectangleSize(20, 20)
    p.scan_step = U.Delta(5, 5)
    p.scan_depth.horizontal, p.scan_depth.vertical = depth, depth
    p.scan_direction = U.Direction(True, True)
    p.abs_threshold = 242
    p.intensity = intensity
    if excl is not None:
        p.exclusions_count = len(excl)
        p.exclusions = excl
    return p


@pytest.mark.parametrize("fmt", FMTS_BYTE)
@pytest.mark.parametrize("intensity", [3, 20])
def test_blackfilter_blobs(cuda_ops, ref_ops, fmt, intensity):
    for seed, (w, h), depth in ((1, (200, 150), 60), (2, (333, 257), 100), (3, (640, 480), 500), (4, (120, 700), 50)):
        img = blobs_image(seed, w, h, fmt)
        excl = (U.Rectangle * 1)(U.rect(w // 4, h // 4, w // 4 + w // 2 - 1, h // 4 + h // 2 - 1))
        for ex in (None, excl):
            p = _bf_params(w, h, ex, intensity, depth)
            a = run_inplace(cuda_ops, "blackfilter", img, fmt, w, C.byref(p))
            b = run_inplace(ref_ops, "blackfilter", img, fmt, w, C.byref(p))
            assert_same(a, b, fmt, w, f"blackfilter seed={seed} I={intensity} excl={ex is not None}")


def test_blackfilter_page(cuda_ops, ref_ops):
    g = synth.gray_page(5, 1240, 1754)
    excl = (U.Rectangle * 1)(U.rect(310, 438, 310 + 620 - 1, 438 + 877 - 1))
    p = _bf_params(1240, 1754, excl)
    a = run_inplace(cuda_ops, "blackfilter", g, U.FMT_GRAY8, 1240, C.byref(p))
    b = run_inplace(ref_ops, "blackfilter", g, U.FMT_GRAY8, 1240, C.byref(p))
    assert_same(a, b, U.FMT_GRAY8, 1240, "blackfilter page")
    assert not np.array_equal(a, g), "dark scan edges should have been filled"


@pytest.mark.parametrize("fmt", FMTS_BYTE)
def test_blurfilter(cuda_ops, ref_ops, fmt):
    for seed, (w, h), (bw, bh), st, inten, dens in ((1, (400, 300), (100, 100), 50, 0.01, 0.002),
                                                    (2, (1000, 1000), (100, 100), 50, 0.01, 0.01),
                                                    (3, (1234, 987), (100, 100), 50, 0.01, 0.004),
                                                    (4, (257, 199), (16, 16), 8, 0.05, 0.05),
                                                    (5, (90, 90), (100, 100), 50, 0.01, 0.01)):
        img = noise_image(seed, w, h, fmt, dark=dens, lo=0, hi=256)
        p = U.BlurfilterParameters(U.RectangleSize(bw, bh), U.Delta(st, st), inten)
        a = run_inplace(cuda_ops, "blurfilter", img, fmt, w, C.byref(p), 229)
        b = run_inplace(ref_ops, "blurfilter", img, fmt, w, C.byref(p), 229)
        assert_same(a, b, fmt, w, f"blurfilter seed={seed}")


@pytest.mark.parametrize("fmt", FMTS_BYTE)
def test_grayfilter(cuda_ops, ref_ops, fmt):
    for seed, (w, h), (sw, sh), st, thr in ((1, (400, 300), (50, 50), 20, 127), (2, (1001, 777), (50, 50), 20, 127),
                                            (3, (640, 480), (30, 20), 10, 100), (4, (260, 200), (50, 50), 20, 200)):
        rng = np.random.Generator(np.random.PCG64(seed))
        g = np.full((h, w), 255, dtype=np.uint8)
        for _ in range(30):   # light-gray blotches (wiped) and a few dark ones (kept)
            x, y = int(rng.integers(0, w - 10)), int(rng.integers(0, h - 10))
            bw, bh = int(rng.integers(10, 120)), int(rng.integers(10, 120))
            g[y:y + bh, x:x + bw] = rng.integers(120, 250) if rng.random() < 0.8 else rng.integers(0, 100)
        sp = rng.random((h, w)) < 0.001
        g[sp] = rng.integers(0, 255, size=int(sp.sum()), dtype=np.uint8)
        img = from_gray(g, fmt)
        if fmt == U.FMT_RGB24:   # decorrelate channels
            v = img[:, :3 * w].reshape(h, w, 3)
            v[..., 1] = np.minimum(255, v[..., 1].astype(int) + rng.integers(0, 6, size=(h, w))).astype(np.uint8)
        p = U.GrayfilterParameters(U.RectangleSize(sw, sh), U.Delta(st, st), thr)
        a = run_inplace(cuda_ops, "grayfilter", img, fmt, w, C.byref(p))
        b = run_inplace(ref_ops, "grayfilter", img, fmt, w, C.byref(p))
        assert_same(a, b, fmt, w, f"grayfilter seed={seed}")


def _mask_params(w, h, horizontal=True, vertical=False, depth=(-1, -1)):
    p = U.MaskDetectionParameters()
    p.scan_size = U.RectangleSize(50, 50)
    p.scan_step = U.Delta(5, 5)
    p.scan_depth.horizontal, p.scan_depth.vertical = depth
    p.scan_direction = U.Direction(horizontal, vertical)
    p.scan_threshold.horizontal = p.scan_threshold.vertical = 0.1
    p.minimum_width = p.minimum_height = 100
    p.maximum_width, p.maximum_height = w, h
    return p


@pytest.mark.parametrize("fmt", [U.FMT_GRAY8, U.FMT_RGB24])
def test_detect_masks(cuda_ops, ref_ops, fmt):
    for idx, (w, h) in enumerate(((1240, 1754), (620, 877), (800, 600))):
        # speckle=0: with noise in the margins the reference's detect_edge() can run
        # off the image and never return (masks.c:88-97)
        g = synth.gray_page(idx, w, h, dark_edges=False, speckle=0)
        img = from_gray(g, fmt)
        pts = (U.Point * 2)(U.Point(w // 2, h // 2), U.Point(w // 3, h // 3))
        for hv in ((True, False), (True, True), (False, True)):
            for depth in ((-1, -1), (200, 300)):
                p = _mask_params(w, h, *hv, depth=depth)
                ma, mb = (U.Rectangle * 2)(), (U.Rectangle * 2)()
                ca = cuda_ops.call("detect_masks", C.byref(himg(img, fmt, w)), C.byref(p), pts, 2, ma)
                cb = ref_ops.call("detect_masks", C.byref(himg(img, fmt, w)), C.byref(p), pts, 2, mb)
                assert ca == cb
                assert [U.rect_tuple(m) for m in ma] == [U.rect_tuple(m) for m in mb], f"masks page{idx} {hv} {depth}"


@pytest.mark.parametrize("fmt", [U.FMT_GRAY8, U.FMT_RGB24])
def test_detect_border(cuda_ops, ref_ops, fmt):
    for idx, (w, h) in enumerate(((620, 877), (800, 600))):
        g = synth.gray_page(idx + 10, w, h, dark_edges=False, speckle=0)
        img = from_gray(g, fmt)
        for outside in (U.rect(0, 0, w - 1, h - 1), U.rect(0, 0, w // 2, h - 1), U.rect(w // 2, 0, w - 1, h - 1),
                        U.rect(-10, -10, w + 10, h + 10)):
            for hv in ((False, True), (True, True), (True, False)):
                p = U.BorderScanParameters()
                p.scan_size = U.RectangleSize(5, 5); p.scan_step = U.Delta(5, 5)
                p.scan_threshold.horizontal = p.scan_threshold.vertical = 5
                p.scan_direction = U.Direction(*hv)
                ba, bb = U.Border(), U.Border()
                cuda_ops.call("detect_border", C.byref(himg(img, fmt, w)), C.byref(p), C.byref(outside), C.byref(ba))
                ref_ops.call("detect_border", C.byref(himg(img, fmt, w)), C.byref(p), C.byref(outside), C.byref(bb))
                assert U.border_tuple(ba) == U.border_tuple(bb), f"border page{idx} {U.rect_tuple(outside)} {hv}"


def _deskew_params(edges=(True, False, True, False), size=1500):
    c = U.default_sheet_config()
    p = c.deskew
    p.scan_edges = U.Edges(*edges)
    p.deskewScanSize = size
    return p


@pytest.mark.parametrize("fmt", [U.FMT_GRAY8, U.FMT_RGB24])
def test_detect_rotation(cuda_ops, ref_ops, fmt):
    """Reference tests/cuda_deskew_test.c:110-111 asks |cpu-cuda| < 1e-6; here: equal."""
    for idx, (w, h) in enumerate(((620, 877), (800, 1000), (1240, 1754))):
        g = synth.gray_page(idx + 20, w, h, dark_edges=False)
        img = from_gray(g, fmt)
        mask = U.rect(int(w * 0.06), 0, int(w * 0.94), h - 1)
        for edges, size in (((True, False, True, False), 1500), ((True, True, True, True), 300), ((False, True, False, False), -1)):
            p = _deskew_params(edges, size)
            ra, rb = C.c_float(), C.c_float()
            cuda_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(ra))
            ref_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(rb))
            assert ra.value == rb.value, f"rotation page{idx} edges={edges}: cuda {ra.value} ref {rb.value}"


@pytest.mark.parametrize("fmt", FMTS_BYTE)
@pytest.mark.parametrize("interp", [U.INTERP_NN, U.INTERP_LINEAR, U.INTERP_CUBIC])
def test_deskew(cuda_ops, ref_ops, fmt, interp):
    """north_star tolerance: +-1 gray level on at most 0.1 % of rotated pixels.
    The kernel is built without FMA and fed the host's sinf/cosf, so it is
    expected (and checked) to be exact; the tolerance is the fallback bar."""
    w, h = 620, 877
    g = synth.gray_page(31, w, h, dark_edges=False)
    img = from_gray(g, fmt)
    for mask, rad in ((U.rect(40, 0, 580, h - 1), 0.0349), (U.rect(100, 100, 400, 500), -0.0610865),
                      (U.rect(-10, 5, 300, 200), 0.0872), (U.rect(0, 0, w - 1, h - 1), -0.001)):
        a = run_inplace(cuda_ops, "deskew", img, fmt, w, C.byref(mask), rad, interp)
        b = run_inplace(ref_ops, "deskew", img, fmt, w, C.byref(mask), rad, interp)
        n = U.bytes_per_row(fmt, w)
        d = np.abs(a[:, :n].astype(int) - b[:, :n].astype(int))
        assert d.max() <= 1 and (d > 0).mean() <= 0.001, f"deskew: max diff {d.max()}, frac {(d > 0).mean()}"
        assert_same(a, b, fmt, w, f"deskew exact {U.rect_tuple(mask)} {rad}")


def test_noisefilter_dense_page_high_intensity(cuda_ops, ref_ops):
    """intensity > 15: every dark pixel is decided in raster order; a page with far more than a
    quarter of its pixels dark must not overflow the list (the reference just processes it)."""
    from util import noise_image, run_inplace, assert_same
    w, h = 320, 200
    img = noise_image(77, w, h, U.FMT_GRAY8, dark=0.6, lo=0, hi=120)
    for inten in (16, 40):
        a = run_inplace(cuda_ops, "noisefilter", img, U.FMT_GRAY8, w, inten, 229)
        b = run_inplace(ref_ops, "noisefilter", img, U.FMT_GRAY8, w, inten, 229)
        assert_same(a, b, U.FMT_GRAY8, w, f"noisefilter intensity {inten} on a dense page")


def test_detect_rotation_beyond_depth_cap(cuda_ops, ref_ops):
    """The column-prefix table of the rotation scan covers the first 256 columns from each scanned
    edge; a page on which a scan line has not reached deskew.c:67's total by then is redone over the
    full width.  Faint ink (the running total grows slowly) and a wide white margin inside the mask
    both need that second pass; the mixed case needs it for one edge only."""
    w, h = 1800, 1200
    base = synth.gray_page(91, w, h, dark_edges=False, box=(0.5, 0.8))       # text box: the middle half of the width
    faint = base.copy()
    faint[base < 128] = 254
    mixed = base.copy()
    mixed[:, :w // 2][base[:, :w // 2] < 128] = 254                           # left half faint, right half dark
    mask = U.rect(0, 0, w - 1, h - 1)                                          # the edges start 450 px away from the text
    for name, g in (("wide margin", base), ("faint", faint), ("mixed", mixed)):
        for size in (1500, 700):
            p = _deskew_params((True, False, True, False), size)
            ra, rb = C.c_float(), C.c_float()
            cuda_ops.call("detect_rotation", C.byref(himg(g, U.FMT_GRAY8, w)), C.byref(mask), C.byref(p), C.byref(ra))
            ref_ops.call("detect_rotation", C.byref(himg(g, U.FMT_GRAY8, w)), C.byref(mask), C.byref(p), C.byref(rb))
            assert ra.value == rb.value, f"{name} size {size}: cuda {ra.value} ref {rb.value}"
