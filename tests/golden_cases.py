"""Seeded cases shared by tests/golden/make_golden.py (which records what the
reference does) and the tests (which hold the oracle restatement and the CUDA
backend to those records)."""
import ctypes as C
import hashlib

import numpy as np

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from util import blobs_image, from_gray, himg, noise_image, run_inplace, visible

SMALL_BOX = (0.60, 0.72)


def result_dict(r):
    return {"deskew_masks": [list(U.rect_tuple(r.deskew_masks[i])) for i in range(r.deskew_mask_count)],
            "rotation": [float(r.rotation[i]) for i in range(r.deskew_mask_count)],
            "center_masks": [list(U.rect_tuple(r.center_masks[i])) for i in range(r.center_mask_count)],
            "borders": [list(U.border_tuple(r.borders[i])) for i in range(r.border_count)]}


def sheet_cases():
    cases = {}
    w, h = 620, 877
    cases["gray_620"] = (U.default_sheet_config(), np.stack([synth.gray_page(i, w, h, box=SMALL_BOX) for i in range(3)]), w, h, U.FMT_GRAY8)
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = 1
    cases["color_620"] = (cfg, np.stack([synth.color_page(i, w, h) for i in range(2)]), w, h, U.FMT_RGB24)
    cfg = U.default_sheet_config()
    cfg.layout = U.LAYOUT_DOUBLE
    cases["double_1754"] = (cfg, np.stack([synth.double_sheet(i, 1754, 1240) for i in range(1)]), 1754, 1240, U.FMT_GRAY8)
    return cases


def _digest(a, fmt, w):
    return hashlib.sha256(visible(a, fmt, w).tobytes()).hexdigest()


def op_cases():
    """name -> fn(ops) -> json-able record of what `ops` produced."""
    def noise(ops):
        out = {}
        for fmt in (U.FMT_GRAY8, U.FMT_RGB24):
            for inten in (2, 4, 7):
                img = noise_image(100 + inten, 300, 200, fmt, dark=0.04)
                out[f"{fmt}_{inten}"] = _digest(run_inplace(ops, "noisefilter", img, fmt, 300, inten, 229), fmt, 300)
        return out

    def black(ops):
        out = {}
        for fmt in (U.FMT_GRAY8, U.FMT_RGB24):
            img = blobs_image(7, 333, 257, fmt)
            p = U.BlackfilterParameters()
            p.scan_size = U.RectangleSize(20, 20); p.scan_step = U.Delta(5, 5)
            p.scan_depth.horizontal = p.scan_depth.vertical = 100
            p.scan_direction = U.Direction(True, True); p.abs_threshold = 242; p.intensity = 20
            out[str(fmt)] = _digest(run_inplace(ops, "blackfilter", img, fmt, 333, C.byref(p)), fmt, 333)
        return out

    def blur_gray(ops):
        out = {}
        img = noise_image(9, 1000, 700, U.FMT_GRAY8, dark=0.004)
        p = U.BlurfilterParameters(U.RectangleSize(100, 100), U.Delta(50, 50), 0.01)
        out["blur"] = _digest(run_inplace(ops, "blurfilter", img, U.FMT_GRAY8, 1000, C.byref(p), 229), U.FMT_GRAY8, 1000)
        rng = np.random.Generator(np.random.PCG64(5))
        g = np.full((480, 640), 255, dtype=np.uint8)
        for _ in range(25):
            x, y = int(rng.integers(0, 600)), int(rng.integers(0, 440))
            g[y:y + int(rng.integers(10, 100)), x:x + int(rng.integers(10, 100))] = rng.integers(60, 250)
        q = U.GrayfilterParameters(U.RectangleSize(50, 50), U.Delta(20, 20), 127)
        out["gray"] = _digest(run_inplace(ops, "grayfilter", g, U.FMT_GRAY8, 640, C.byref(q)), U.FMT_GRAY8, 640)
        return out

    def detect(ops):
        w, h = 800, 1000
        g = synth.gray_page(21, w, h, dark_edges=False, speckle=0)
        cfg = U.default_sheet_config()
        mp = cfg.mask_detection
        mp.maximum_width, mp.maximum_height = w, h
        pts = (U.Point * 1)(U.Point(w // 2, h // 2))
        m = (U.Rectangle * 1)()
        ops.call("detect_masks", C.byref(himg(g, U.FMT_GRAY8, w)), C.byref(mp), pts, 1, m)
        rot = C.c_float()
        ops.call("detect_rotation", C.byref(himg(g, U.FMT_GRAY8, w)), C.byref(m[0]), C.byref(cfg.deskew), C.byref(rot))
        b = U.Border()
        full = U.rect(0, 0, w - 1, h - 1)
        ops.call("detect_border", C.byref(himg(g, U.FMT_GRAY8, w)), C.byref(cfg.border_scan), C.byref(full), C.byref(b))
        d = run_inplace(ops, "deskew", g, U.FMT_GRAY8, w, C.byref(m[0]), rot.value, U.INTERP_CUBIC)
        return {"mask": list(U.rect_tuple(m[0])), "rotation": float(rot.value), "border": list(U.border_tuple(b)),
                "deskew": _digest(d, U.FMT_GRAY8, w)}

    return {"noisefilter": noise, "blackfilter": black, "blur_gray": blur_gray, "detect": detect}


def c1_config():
    """The options of the reference's test_c1_mask_border_scan_fixture
    (tests/unpaper_tests.py:575-595)."""
    cfg = U.default_sheet_config()
    cfg.no_deskew = cfg.no_blackfilter = cfg.no_noisefilter = cfg.no_blurfilter = cfg.no_grayfilter = 1
    cfg.no_mask_center = 1
    cfg.mask_detection.scan_direction = U.Direction(True, True)
    cfg.mask_detection.scan_threshold.horizontal = cfg.mask_detection.scan_threshold.vertical = 0.8
    cfg.mask_detection.minimum_width = cfg.mask_detection.minimum_height = 1
    cfg.border_scan.scan_direction = U.Direction(True, True)
    cfg.pre_wipe_count = 1
    cfg.pre_wipes[0] = U.rect(0, 0, 9, 9)
    cfg.pre_border = U.Border(2, 2, 2, 2)
    return cfg

