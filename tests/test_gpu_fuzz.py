"""Seeded parameter fuzz, CUDA vs reference CPU backend — bit-exact.

The fixed cases of test_gpu_filters.py pin the known quirks; this file walks the
parameter space around them (sizes with and without 16-byte aligned rows — which
switches the flood fill's shared-memory column strip and the noise filter's
bit-plane path on and off —, scan sizes/steps/directions, intensities, thresholds)."""
import ctypes as C

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from util import assert_same, blobs_image, from_gray, noise_image, run_inplace

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.Generator(np.random.PCG64(0xF022 + seed))


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_blackfilter(cuda_ops, ref_ops, seed):
    r = _rng(seed)
    w = int(r.choice([96, 160, 208, 250, 333, 400, 512, 641]))
    h = int(r.integers(80, 420))
    fmt = U.FMT_GRAY8 if seed % 4 else U.FMT_RGB24
    img = blobs_image(100 + seed, w, h, fmt, nblobs=int(r.integers(2, 10)))
    if seed % 3 == 0:   # a dark scan edge on one side, the page-scale case in small
        g = img[:, :w] if fmt == U.FMT_GRAY8 else None
        if g is not None:
            g[:, :int(r.integers(8, 40))] = int(r.integers(0, 30))
    p = U.BlackfilterParameters()
    s = int(r.choice([8, 12, 20, 25]))
    p.scan_size = U.RectangleSize(s, int(r.choice([s, s + 5])))
    st = int(r.choice([3, 5, 10]))
    p.scan_step = U.Delta(st, st)
    p.scan_depth.horizontal, p.scan_depth.vertical = int(r.integers(30, 600)), int(r.integers(30, 600))
    p.scan_direction = U.Direction(bool(seed % 5), bool((seed + 1) % 3))
    p.abs_threshold = int(r.choice([200, 230, 242]))
    p.intensity = int(r.choice([1, 3, 8, 20, 40]))
    a = run_inplace(cuda_ops, "blackfilter", img, fmt, w, C.byref(p))
    b = run_inplace(ref_ops, "blackfilter", img, fmt, w, C.byref(p))
    assert_same(a, b, fmt, w, f"fuzz blackfilter seed={seed} {w}x{h}")


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_noisefilter(cuda_ops, ref_ops, seed):
    r = _rng(seed)
    w = int(r.choice([64, 100, 192, 193, 256, 300, 448, 500]))
    h = int(r.integers(40, 300))
    fmt = [U.FMT_GRAY8, U.FMT_GRAY8, U.FMT_RGB24, U.FMT_MONOWHITE][seed % 4]
    img = noise_image(200 + seed, w, h, fmt, dark=float(r.choice([0.005, 0.02, 0.08, 0.2])), lo=0, hi=256)
    intensity = int(r.choice([1, 2, 3, 4, 5, 7, 8, 12]))
    white = int(r.choice([200, 229, 250]))
    a = run_inplace(cuda_ops, "noisefilter", img, fmt, w, intensity, white)
    b = run_inplace(ref_ops, "noisefilter", img, fmt, w, intensity, white)
    assert_same(a, b, fmt, w, f"fuzz noisefilter seed={seed} {w}x{h} I={intensity}")


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_gray_blur(cuda_ops, ref_ops, seed):
    r = _rng(seed)
    w, h = int(r.choice([200, 256, 333, 480])), int(r.integers(120, 400))
    g = np.full((h, w), 255, dtype=np.uint8)
    for _ in range(int(r.integers(5, 40))):
        x, y = int(r.integers(0, w - 10)), int(r.integers(0, h - 10))
        bw, bh = int(r.integers(5, 90)), int(r.integers(5, 90))
        g[y:y + bh, x:x + bw] = r.integers(120, 250) if r.random() < 0.8 else r.integers(0, 100)
    sp = r.random((h, w)) < 0.002
    g[sp] = r.integers(0, 255, size=int(sp.sum()), dtype=np.uint8)
    fmt = U.FMT_GRAY8 if seed % 3 else U.FMT_RGB24
    img = from_gray(g, fmt)
    size = int(r.choice([20, 30, 50]))
    step = int(r.choice([5, 10, 20, 25]))
    gp = U.GrayfilterParameters(U.RectangleSize(size, int(r.choice([size, size + 10]))), U.Delta(step, step),
                                int(r.choice([100, 127, 200])))
    a = run_inplace(cuda_ops, "grayfilter", img, fmt, w, C.byref(gp))
    b = run_inplace(ref_ops, "grayfilter", img, fmt, w, C.byref(gp))
    assert_same(a, b, fmt, w, f"fuzz grayfilter seed={seed}")
    bs = int(r.choice([16, 40, 100]))
    bp = U.BlurfilterParameters(U.RectangleSize(bs, bs), U.Delta(bs // 2, bs // 2), float(r.choice([0.01, 0.05, 0.2])))
    a = run_inplace(cuda_ops, "blurfilter", img, fmt, w, C.byref(bp), 229)
    b = run_inplace(ref_ops, "blurfilter", img, fmt, w, C.byref(bp), 229)
    assert_same(a, b, fmt, w, f"fuzz blurfilter seed={seed}")


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_deskew(cuda_ops, ref_ops, seed):
    r = _rng(seed)
    w, h = int(r.choice([160, 257, 320, 400])), int(r.integers(120, 360))
    fmt = [U.FMT_GRAY8, U.FMT_RGB24, U.FMT_GRAY8, U.FMT_Y400A][seed % 4]
    img = noise_image(300 + seed, w, h, fmt, dark=0.3, lo=0, hi=256)
    x0, y0 = int(r.integers(0, w // 3)), int(r.integers(0, h // 3))
    mask = U.rect(x0, y0, int(r.integers(w // 2, w)), int(r.integers(h // 2, h)))
    rad = float(np.float32(r.uniform(-0.12, 0.12)))
    for interp in (0, 1, 2):
        a = run_inplace(cuda_ops, "deskew", img, fmt, w, C.byref(mask), C.c_float(rad), interp)
        b = run_inplace(ref_ops, "deskew", img, fmt, w, C.byref(mask), C.c_float(rad), interp)
        assert_same(a, b, fmt, w, f"fuzz deskew seed={seed} interp={interp} rad={rad}")


def _page(seed, w, h, fmt):
    # clean margins: with noise in the margins the reference's detect_edge() can run off
    # the image and never return (masks.c:88-97)
    from unpaper_gpu_b200 import synth
    return from_gray(synth.gray_page(seed, w, h, dark_edges=False, speckle=0), fmt)


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_detect_rotation(cuda_ops, ref_ops, seed):
    from util import himg
    r = _rng(seed)
    w, h = int(r.choice([400, 620, 800])), int(r.choice([500, 877, 1000]))
    fmt = U.FMT_GRAY8 if seed % 3 else U.FMT_RGB24
    img = _page(400 + seed, w, h, fmt)
    p = U.default_sheet_config().deskew
    p.scan_edges = U.Edges(*[bool(x) for x in r.integers(0, 2, 4)])
    if not any((p.scan_edges.left, p.scan_edges.top, p.scan_edges.right, p.scan_edges.bottom)):
        p.scan_edges = U.Edges(True, False, False, False)
    p.deskewScanSize = int(r.choice([-1, 100, 300, 1500]))
    p.deskewScanDepth = float(r.choice([0.1, 0.5, 0.9]))
    p.deskewScanRangeRad = float(np.float32(np.deg2rad(float(r.choice([2.0, 5.0, 9.0])))))
    p.deskewScanStepRad = float(np.float32(np.deg2rad(float(r.choice([0.1, 0.25, 0.5])))))
    p.deskewScanDeviationRad = float(np.float32(np.deg2rad(float(r.choice([0.2, 1.0, 5.0])))))
    x0, x1 = int(r.integers(0, w // 4)), int(r.integers(3 * w // 4, w))
    y0, y1 = int(r.integers(0, h // 4)), int(r.integers(3 * h // 4, h))
    mask = U.rect(x0, y0, x1, y1)
    ra, rb = C.c_float(), C.c_float()
    cuda_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(ra))
    ref_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(rb))
    assert ra.value == rb.value, f"fuzz rotation seed={seed}: cuda {ra.value} ref {rb.value}"


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_detect_border_and_masks(cuda_ops, ref_ops, seed):
    from util import himg
    r = _rng(seed)
    w, h = int(r.choice([500, 620, 800])), int(r.choice([600, 877]))
    fmt = U.FMT_GRAY8 if seed % 3 else U.FMT_RGB24
    img = _page(500 + seed, w, h, fmt)
    bp = U.BorderScanParameters()
    s = int(r.choice([3, 5, 9]))
    bp.scan_size = U.RectangleSize(s, s)
    bp.scan_step = U.Delta(int(r.choice([2, 5, 7])), int(r.choice([2, 5, 7])))
    bp.scan_threshold.horizontal = int(r.choice([1, 5, 20]))
    bp.scan_threshold.vertical = int(r.choice([1, 5, 20]))
    bp.scan_direction = U.Direction(bool(seed % 2), True)
    outside = U.rect(int(r.integers(-5, 40)), int(r.integers(-5, 40)), w - 1 - int(r.integers(-5, 40)), h - 1 - int(r.integers(-5, 40)))
    ba, bb = U.Border(), U.Border()
    cuda_ops.call("detect_border", C.byref(himg(img, fmt, w)), C.byref(bp), C.byref(outside), C.byref(ba))
    ref_ops.call("detect_border", C.byref(himg(img, fmt, w)), C.byref(bp), C.byref(outside), C.byref(bb))
    assert U.border_tuple(ba) == U.border_tuple(bb), f"fuzz border seed={seed}"
    mp = U.MaskDetectionParameters()
    ms = int(r.choice([30, 50, 80]))
    mp.scan_size = U.RectangleSize(ms, ms)
    mp.scan_step = U.Delta(int(r.choice([3, 5, 10])), int(r.choice([3, 5, 10])))
    mp.scan_depth.horizontal, mp.scan_depth.vertical = (-1, -1) if seed % 2 else (int(r.integers(100, 400)), int(r.integers(100, 400)))
    mp.scan_direction = U.Direction(True, bool(seed % 3 == 0))
    mp.scan_threshold.horizontal = mp.scan_threshold.vertical = float(r.choice([0.05, 0.1, 0.2]))
    mp.minimum_width = mp.minimum_height = int(r.choice([50, 100, 300]))
    mp.maximum_width, mp.maximum_height = int(r.choice([w, w // 2])), h
    pts = (U.Point * 2)(U.Point(w // 2, h // 2), U.Point(int(w * 0.4), int(h * 0.45)))
    ma, mb = (U.Rectangle * 2)(), (U.Rectangle * 2)()
    ca = cuda_ops.call("detect_masks", C.byref(himg(img, fmt, w)), C.byref(mp), pts, 2, ma)
    cb = ref_ops.call("detect_masks", C.byref(himg(img, fmt, w)), C.byref(mp), pts, 2, mb)
    assert ca == cb
    assert [U.rect_tuple(m) for m in ma] == [U.rect_tuple(m) for m in mb], f"fuzz masks seed={seed}"


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_moves(cuda_ops, ref_ops, seed):
    r = _rng(seed)
    w, h = int(r.choice([160, 256, 333])), int(r.integers(100, 300))
    fmt = [U.FMT_GRAY8, U.FMT_RGB24, U.FMT_MONOWHITE][seed % 3]
    img = noise_image(600 + seed, w, h, fmt, dark=0.3)
    ax0, ay0 = int(r.integers(-10, w // 2)), int(r.integers(-10, h // 2))
    area = U.rect(ax0, ay0, ax0 + int(r.integers(10, w // 2)), ay0 + int(r.integers(10, h // 2)))
    center = U.Point(int(r.integers(0, w)), int(r.integers(0, h)))
    a = run_inplace(cuda_ops, "center_mask", img, fmt, w, center, C.byref(area))
    b = run_inplace(ref_ops, "center_mask", img, fmt, w, center, C.byref(area))
    assert_same(a, b, fmt, w, f"fuzz center_mask seed={seed}")
    p = U.MaskAlignmentParameters()
    p.alignment = U.Edges(*[bool(x) for x in r.integers(0, 2, 4)])
    p.margin = U.Delta(int(r.integers(0, 20)), int(r.integers(0, 20)))
    outside = U.rect(int(r.integers(0, 20)), int(r.integers(0, 20)), w - 1 - int(r.integers(0, 20)), h - 1 - int(r.integers(0, 20)))
    a = run_inplace(cuda_ops, "align_mask", img, fmt, w, C.byref(area), C.byref(outside), C.byref(p))
    b = run_inplace(ref_ops, "align_mask", img, fmt, w, C.byref(area), C.byref(outside), C.byref(p))
    assert_same(a, b, fmt, w, f"fuzz align_mask seed={seed}")


def test_detect_rotation_steep_range(cuda_ops, ref_ops):
    """A scan range steep enough (30 degrees over 1500 rows: > 700 column changes per
    line) that the warp-per-angle kernel's shared-memory run table does not apply and
    the one-CTA-per-angle form runs instead."""
    from util import himg
    w, h = 1240, 1754
    fmt = U.FMT_GRAY8
    img = _page(700, w, h, fmt)
    p = U.default_sheet_config().deskew
    p.deskewScanRangeRad = float(np.float32(np.deg2rad(30.0)))
    p.deskewScanStepRad = float(np.float32(np.deg2rad(0.5)))
    p.deskewScanDeviationRad = float(np.float32(np.deg2rad(5.0)))
    mask = U.rect(int(w * 0.06), 0, int(w * 0.94), h - 1)
    ra, rb = C.c_float(), C.c_float()
    cuda_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(ra))
    ref_ops.call("detect_rotation", C.byref(himg(img, fmt, w)), C.byref(mask), C.byref(p), C.byref(rb))
    assert ra.value == rb.value, f"cuda {ra.value} ref {rb.value}"
