"""CPU-only: pins the oracle restatement (oracle/liboracle.so).

 * against the unmodified reference compiled into oracle/_ref (when present —
   it is in the build container and it travels to the GPU box as a prebuilt
   .so): every op bit-exact on seeded inputs, and the whole process_sheet()
   pipeline on small synthetic sheets;
 * against tests/golden/golden_vectors.json, which records what the reference
   produced (incl. its own goldenA1 end-to-end case) — works without oracle/_ref.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from oracle import checker  # test infrastructure: the CPU checkers
import golden_cases as G
import test_gpu_blit as B
import test_gpu_filters as F
from util import FMTS_ALL, FMTS_BYTE

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_vectors.json")))


@pytest.fixture(scope="module")
def orc_lib():
    lib = checker.load_oracle()
    if lib is None or not hasattr(lib, "orc_process_sheets"):
        pytest.fail("oracle/liboracle.so not built: run __graft_entry__.build()")
    return lib


@pytest.fixture(scope="module")
def orc_ops(orc_lib):
    return U.HostOps(orc_lib, "orc_host_")


def test_golden_a1_record_is_sane():
    """oracle/_ref reproduced the reference's goldenA1.pbm when the vectors were made."""
    a1 = GOLD["A1"]
    assert a1 is not None and a1["golden_diff_ratio_thr170"] < 1e-4
    assert a1["result"]["center_masks"] == [[670, 0, 2265, 3506]]
    assert a1["result"]["borders"] == [[0, 315, 1, 741]]


@pytest.mark.parametrize("name", sorted(GOLD["ops"].keys()))
def test_oracle_ops_match_golden(orc_ops, name):
    got = G.op_cases()[name](orc_ops)
    assert got == GOLD["ops"][name]


@pytest.mark.parametrize("name", ["gray_620", "color_620"])
def test_oracle_sheets_match_golden(orc_lib, name):
    cfg, pages, w, h, fmt = G.sheet_cases()[name]
    out, res = checker.process_sheets_cpu(orc_lib, "orc_", cfg, pages, w, h, fmt, threads=4)
    want = GOLD["sheets"][name]
    assert [G.result_dict(r) for r in res] == want["results"]
    assert [hashlib.sha256(x.tobytes()).hexdigest() for x in out] == want["output_sha256"]


def test_ref_still_matches_golden(ref_lib):
    """The prebuilt reference library agrees with the committed vectors."""
    cfg, pages, w, h, fmt = G.sheet_cases()["gray_620"]
    out, res = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages[:1], w, h, fmt)
    assert G.result_dict(res[0]) == GOLD["sheets"]["gray_620"]["results"][0]
    assert hashlib.sha256(out[0].tobytes()).hexdigest() == GOLD["sheets"]["gray_620"]["output_sha256"][0]


# ---- op-by-op against the live reference (same bodies as the GPU parity tests) ----

@pytest.mark.parametrize("fmt", FMTS_ALL)
def test_orc_blit_ops(orc_ops, ref_ops, fmt):
    B.test_wipe_rectangle(orc_ops, ref_ops, fmt, 37, 29)
    B.test_mirror(orc_ops, ref_ops, fmt, 37, 29)
    B.test_flip_rotate_90(orc_ops, ref_ops, fmt)
    B.test_shift(orc_ops, ref_ops, fmt)
    B.test_apply_masks_wipes_border(orc_ops, ref_ops, fmt)
    B.test_center_and_align_mask(orc_ops, ref_ops, fmt)


def test_orc_copy_center(orc_ops, ref_ops):
    for s, d in ((U.FMT_GRAY8, U.FMT_RGB24), (U.FMT_RGB24, U.FMT_GRAY8), (U.FMT_Y400A, U.FMT_Y400A),
                 (U.FMT_MONOWHITE, U.FMT_RGB24), (U.FMT_RGB24, U.FMT_MONOBLACK)):
        B.test_copy_rectangle(orc_ops, ref_ops, s, d)
    B.test_center_image(orc_ops, ref_ops, U.FMT_RGB24)


@pytest.mark.parametrize("interp", [U.INTERP_NN, U.INTERP_LINEAR, U.INTERP_CUBIC])
def test_orc_stretch_deskew(orc_ops, ref_ops, interp):
    B.test_stretch_resize(orc_ops, ref_ops, U.FMT_RGB24, interp)
    F.test_deskew(orc_ops, ref_ops, U.FMT_GRAY8, interp)


@pytest.mark.parametrize("fmt", FMTS_BYTE)
def test_orc_filters(orc_ops, ref_ops, fmt):
    F.test_noisefilter_random(orc_ops, ref_ops, fmt, 4)
    F.test_noisefilter_random(orc_ops, ref_ops, fmt, 7)
    F.test_blackfilter_blobs(orc_ops, ref_ops, fmt, 20)
    F.test_blurfilter(orc_ops, ref_ops, fmt)
    F.test_grayfilter(orc_ops, ref_ops, fmt)


def test_orc_detectors(orc_ops, ref_ops):
    F.test_detect_masks(orc_ops, ref_ops, U.FMT_GRAY8)
    F.test_detect_border(orc_ops, ref_ops, U.FMT_RGB24)
    F.test_detect_rotation(orc_ops, ref_ops, U.FMT_GRAY8)


def test_orc_sheet_vs_ref(orc_lib, ref_lib):
    cfg, pages, w, h, fmt = G.sheet_cases()["double_1754"]
    a, ra = checker.process_sheets_cpu(orc_lib, "orc_", cfg, pages, w, h, fmt)
    b, rb = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, fmt)
    assert [G.result_dict(r) for r in ra] == [G.result_dict(r) for r in rb]
    assert np.array_equal(a, b)


def test_orc_output_conversion_vs_numpy(orc_ops):
    """saveImage()'s conversion (file.c:211-259) restated in oracle_sheet.c against a
    numpy statement of the same rule: bit = gray < abs_black_threshold, MSB first,
    tail bits of the last byte clear."""
    import ctypes as C
    from util import himg, linesize, noise_image
    w, h = 203, 9
    for abt in (170, 84):
        src = noise_image(11, w, h, U.FMT_GRAY8, dark=0.4)
        d = np.full((h, linesize(U.FMT_MONOWHITE, w)), 0x5A, dtype=np.uint8)
        orc_ops.call("convert_format", C.byref(himg(src, U.FMT_GRAY8, w, abt=abt)), C.byref(himg(d, U.FMT_MONOWHITE, w, abt=abt)))
        want = np.packbits(src[:, :w] < abt, axis=1)
        assert np.array_equal(d[:, :want.shape[1]], want)
        rgb = noise_image(12, w, h, U.FMT_RGB24, dark=0.4)
        orc_ops.call("convert_format", C.byref(himg(rgb, U.FMT_RGB24, w, abt=abt)), C.byref(himg(d, U.FMT_MONOWHITE, w, abt=abt)))
        g = rgb[:, :3 * w].reshape(h, w, 3).astype(int).sum(2) // 3
        assert np.array_equal(d[:, :want.shape[1]], np.packbits(g < abt, axis=1))
        mb = noise_image(13, w, h, U.FMT_MONOBLACK, dark=0.4)
        orc_ops.call("convert_format", C.byref(himg(mb, U.FMT_MONOBLACK, w, abt=abt)), C.byref(himg(d, U.FMT_MONOWHITE, w, abt=abt)))
        assert np.array_equal(d[:, :want.shape[1]], mb[:, :want.shape[1]] ^ 0xFF)


def test_orc_output_conversion_vs_reference_saveimage(orc_ops, ref_lib):
    """The restated conversion against the reference's own saveImage() (file.c:186-262,
    compiled unmodified into oracle/_ref): every source/output format pair."""
    import ctypes as C
    from util import himg, linesize, noise_image
    pairs = [(U.FMT_GRAY8, U.FMT_MONOWHITE), (U.FMT_RGB24, U.FMT_MONOWHITE), (U.FMT_MONOBLACK, U.FMT_MONOWHITE),
             (U.FMT_Y400A, U.FMT_MONOWHITE), (U.FMT_RGB24, U.FMT_GRAY8), (U.FMT_GRAY8, U.FMT_RGB24),
             (U.FMT_MONOWHITE, U.FMT_GRAY8), (U.FMT_MONOBLACK, U.FMT_RGB24), (U.FMT_Y400A, U.FMT_GRAY8),
             (U.FMT_GRAY8, U.FMT_GRAY8), (U.FMT_MONOWHITE, U.FMT_MONOWHITE)]
    for sfmt, dfmt in pairs:
        for w, h in ((37, 29), (203, 77)):
            for abt in (170, 84):
                src = noise_image(11, w, h, sfmt, dark=0.4)
                d = np.zeros((h, linesize(dfmt, w)), dtype=np.uint8)
                orc_ops.call("convert_format", C.byref(himg(src, sfmt, w, abt=abt)), C.byref(himg(d, dfmt, w, abt=abt)))
                rfmt, rw, rh, ref = checker.save_image_cpu(ref_lib, himg(src, sfmt, w, abt=abt), dfmt)
                assert (rfmt, rw, rh) == (dfmt, w, h)
                a, b = d[:, :U.bytes_per_row(dfmt, w)].copy(), ref.copy()
                if dfmt == U.FMT_MONOWHITE and sfmt not in (U.FMT_GRAY8, U.FMT_RGB24, U.FMT_MONOBLACK) and w % 8:
                    keep = (0xFF << (8 - w % 8)) & 0xFF      # generic branch: tail bits unspecified
                    a[:, -1] &= keep
                    b[:, -1] &= keep
                assert np.array_equal(a, b), (sfmt, dfmt, w, h, abt)


def _c1_fixture():
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_fixture.npz"))
    return f["page"], f["golden"]


def test_reference_goldens_c1_f3_e1_records():
    """When the vectors were made, oracle/_ref reproduced the reference's goldenC1.ppm and
    goldenF.pbm exactly and the six goldenE1 pages to < 1e-4 (tests/unpaper_tests.py:568-599, :763-810)."""
    assert GOLD["C1"]["equals_reference_golden"] is True
    assert GOLD["F3"]["differing_pixels_vs_goldenF"] == 0
    assert len(GOLD["E1"]["golden_diff_ratio"]) == 6 and max(GOLD["E1"]["golden_diff_ratio"]) < 1e-4


@pytest.mark.parametrize("which", ["orc", "ref"])
def test_golden_c1_fixture_exact(orc_lib, which):
    """The reference's own exact golden for this path (mask/border scan + pre-wipe/border,
    unpaper_tests.py:568-599) against the restatement and, when built, the reference."""
    import golden_cases as G
    page, golden = _c1_fixture()
    h, w, _ = page.shape
    if which == "ref":
        lib, prefix = checker.load_ref(), "ref_"
        if lib is None:
            pytest.skip("oracle/_ref not built")
    else:
        lib, prefix = orc_lib, "orc_"
    out, res = checker.process_sheets_cpu(lib, prefix, G.c1_config(), page.reshape(1, h, 3 * w), w, h, U.FMT_RGB24)
    assert np.array_equal(out[0].reshape(h, w, 3), golden)
    assert G.result_dict(res[0]) == GOLD["C1"]["result"]


def _small_pages(seed0, n, w=620, h=877):
    return np.stack([synth.gray_page(seed0 + i, w, h, box=(0.60, 0.72)) for i in range(n)])


@pytest.mark.parametrize("case", ["mirror_shift", "two_pages", "wipes_borders", "mono_pages", "stage_switches"])
def test_orc_sheet_vs_ref_configurations(orc_lib, ref_lib, case):
    """The restatement against the unmodified reference over the configuration space the
    engine tests use: geometry options, two pages per sheet, static rectangles, 1-bit pages,
    stage switches — decisions and every output byte."""
    w, h, fmt = 620, 877, U.FMT_GRAY8
    cfg = U.default_sheet_config()
    pages = _small_pages(300, 2)
    if case == "mirror_shift":
        cfg.pre_mirror, cfg.pre_shift = U.Direction(True, False), U.Delta(7, -5)
        cfg.post_mirror, cfg.post_shift = U.Direction(False, True), U.Delta(-3, 11)
    elif case == "two_pages":
        cfg.input_count, cfg.layout = 2, U.LAYOUT_DOUBLE
        pages = _small_pages(310, 4)
    elif case == "wipes_borders":
        cfg.pre_wipe_count = 1; cfg.pre_wipes[0] = U.rect(100, 120, 160, 170)
        cfg.wipe_count = 2; cfg.wipes[0] = U.rect(300, 400, 340, 460); cfg.wipes[1] = U.rect(-5, 800, 50, 900)
        cfg.post_wipe_count = 1; cfg.post_wipes[0] = U.rect(500, 50, 619, 90)
        cfg.pre_border = U.Border(3, 4, 5, 6); cfg.border = U.Border(10, 0, 0, 12); cfg.post_border = U.Border(0, 7, 8, 0)
        cfg.pre_mask_count = 1; cfg.pre_masks[0] = U.rect(20, 20, 600, 860)
        cfg.noisefilter_intensity = 6
        cfg.point_count = 1; cfg.points[0] = U.Point(300, 430)
        cfg.mask_color = U.Pixel(200, 200, 200); cfg.sheet_background = U.Pixel(250, 250, 250)
    elif case == "mono_pages":
        w, h, fmt = 624, 880, U.FMT_MONOWHITE
        pages = np.stack([np.packbits(synth.gray_page(320 + i, w, h, box=(0.60, 0.72)) < 128, axis=1) for i in range(2)])
    elif case == "stage_switches":
        cfg.no_mask_center = cfg.no_border_align = cfg.no_grayfilter = 1
    a, ra = checker.process_sheets_cpu(orc_lib, "orc_", cfg, pages, w, h, fmt, threads=2)
    b, rb = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, fmt, threads=2)
    assert [G.result_dict(r) for r in ra] == [G.result_dict(r) for r in rb]
    assert np.array_equal(a, b)


def test_orc_ops_fuzz_vs_ref(orc_ops, ref_ops):
    """The seeded parameter fuzz of test_gpu_fuzz.py, restatement against the live reference."""
    import test_gpu_fuzz as Z
    for seed in range(16):
        Z.test_fuzz_blackfilter(orc_ops, ref_ops, seed)
        Z.test_fuzz_noisefilter(orc_ops, ref_ops, seed)
    for seed in range(10):
        Z.test_fuzz_gray_blur(orc_ops, ref_ops, seed)
        Z.test_fuzz_deskew(orc_ops, ref_ops, seed)
        Z.test_fuzz_detect_border_and_masks(orc_ops, ref_ops, seed)
        Z.test_fuzz_moves(orc_ops, ref_ops, seed)
    for seed in range(12):
        Z.test_fuzz_detect_rotation(orc_ops, ref_ops, seed)
