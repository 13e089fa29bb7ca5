"""Sheet engine (CUDA) vs the reference's own process_sheet() on the CPU backend:
every decision bit-identical, every pixel identical."""
import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from oracle import checker  # test infrastructure: the CPU checkers

pytestmark = pytest.mark.gpu

# small pages keep a wide white gutter so that the reference's detect_edge() terminates
SMALL_BOX = (0.60, 0.72)


def _compare(cfg, pages, w, h, fmt, ref_lib, group=4, lanes=2):
    from unpaper_gpu_b200.lib import Engine
    eng = Engine(cfg, w, h, fmt, group_pages=group, lanes=lanes)
    out, res = eng.process_numpy(pages)
    size = (eng.sheet_w, eng.sheet_h)
    # the device-resident entry point (sheets rendered straight into the caller's device buffer when
    # the layout allows it) returns the same bytes as the host-buffer one
    import torch
    d_in = torch.from_numpy(np.ascontiguousarray(pages).reshape(-1)).cuda()
    d_out = torch.zeros(out.size, dtype=torch.uint8, device="cuda")
    eng.process_ptr(d_in.data_ptr(), d_out.data_ptr(), len(res), False, None)
    assert np.array_equal(d_out.cpu().numpy().reshape(out.shape), out), "device-resident output differs from the host-buffer output"
    eng.close()
    rout, rres = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, fmt, threads=8, out_size=size)
    for i, (a, b) in enumerate(zip(res, rres)):
        assert a.status == 0 and b.status == 0
        assert a.deskew_mask_count == b.deskew_mask_count, f"sheet {i}"
        for k in range(a.deskew_mask_count):
            assert U.rect_tuple(a.deskew_masks[k]) == U.rect_tuple(b.deskew_masks[k]), f"sheet {i} deskew mask {k}"
            assert a.rotation[k] == b.rotation[k], f"sheet {i} rotation {k}: {a.rotation[k]} vs {b.rotation[k]}"
        assert a.center_mask_count == b.center_mask_count
        for k in range(a.center_mask_count):
            assert U.rect_tuple(a.center_masks[k]) == U.rect_tuple(b.center_masks[k]), f"sheet {i} center mask {k}"
        assert a.border_count == b.border_count
        for k in range(a.border_count):
            assert U.border_tuple(a.borders[k]) == U.border_tuple(b.borders[k]), f"sheet {i} border {k}"
    diff = out != rout
    assert not diff.any(), f"{int(diff.sum())} differing bytes; per sheet {diff.reshape(len(res), -1).sum(axis=1)}"
    return out, res


def test_engine_gray_small(ref_lib):
    w, h = 620, 877
    pages = np.stack([synth.gray_page(i, w, h, box=SMALL_BOX) for i in range(6)])
    _compare(U.default_sheet_config(), pages, w, h, U.FMT_GRAY8, ref_lib, group=4, lanes=2)


def test_engine_gray_half_scale(ref_lib):
    w, h = 1240, 1754
    pages = np.stack([synth.gray_page(40 + i, w, h) for i in range(3)])
    _compare(U.default_sheet_config(), pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)


def test_engine_rgb_small(ref_lib):
    w, h = 620, 877
    pages = np.stack([synth.color_page(i, w, h) for i in range(3)])
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = 1   # BASELINE config 3: gray/blur filters + cubic deskew
    _compare(cfg, pages, w, h, U.FMT_RGB24, ref_lib, group=2, lanes=1)


def test_engine_double_layout(ref_lib):
    w, h = 1754, 1240
    pages = np.stack([synth.double_sheet(i, w, h) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.layout = U.LAYOUT_DOUBLE
    _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=1)


def test_engine_stage_switches(ref_lib):
    w, h = 620, 877
    pages = np.stack([synth.gray_page(70 + i, w, h, box=SMALL_BOX) for i in range(2)])
    for flags in (("no_deskew",), ("no_mask_center", "no_border_align"), ("no_mask_scan",), ("no_border_scan", "no_grayfilter")):
        cfg = U.default_sheet_config()
        for f in flags:
            setattr(cfg, f, 1)
        _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=1)


def _textured_page(index, w, h, box=SMALL_BOX):
    """A page with a light paper texture (a quarter of the background pixels at 253 / 254): no tile of the
    rotation can be skipped as white and most 4x4 neighbourhoods are not constant.  (A heavier texture keeps
    the reference's detect_edge() from ever seeing a light bar: masks.c:88-97 never returns then.)"""
    page = synth.gray_page(index, w, h, speckle=0, dark_edges=False, box=box)
    rng = np.random.Generator(np.random.PCG64(977 + index))
    tex = rng.integers(1, 3, size=page.shape, dtype=np.uint8) * (rng.random(page.shape) < 0.25)
    return np.where(page == 255, 255 - tex, page).astype(np.uint8)


@pytest.mark.parametrize("w,h,box", [(701, 903, SMALL_BOX), (1000, 1300, SMALL_BOX), (1111, 1403, SMALL_BOX), (800, 1000, (0.9, 0.72))])
def test_engine_deskew_textured_pages(ref_lib, w, h, box):
    """Sheet rotation (GRAY8, cubic) with every pixel computed, on widths that are no multiple of the
    rotation tile (128 x 8), of 16 or of 4: tiles that straddle the mask's edges and the sheet's right edge.
    The masks found here run from the top to the bottom row of the sheet (the last one covers the whole sheet
    and one column more), so the rotated rectangle's source footprint leaves the image (reads outside = white):
    the tiles along those edges take the general per-pixel path while the interior keeps the fast one."""
    pages = np.stack([_textured_page(300 + i, w, h, box=box) for i in range(3)])
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = cfg.no_blurfilter = cfg.no_grayfilter = 1
    out, res = _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=3, lanes=1)
    assert any(r.rotation[0] != 0.0 for r in res)


def test_engine_full_a4(ref_lib):
    """BASELINE config 2 at full size, one sheet (the reference needs ~15 s)."""
    w, h = synth.A4_W, synth.A4_H
    pages = synth.gray_page(0, w, h)[None]
    _compare(U.default_sheet_config(), pages, w, h, U.FMT_GRAY8, ref_lib, group=1, lanes=1)


def test_engine_idempotent_and_deterministic():
    """Size-independent properties at full size: same input twice -> same bytes;
    a processed page run again detects rotation 0."""
    from unpaper_gpu_b200.lib import Engine
    w, h = synth.A4_W, synth.A4_H
    pages = np.stack([synth.gray_page(i, w, h) for i in range(4)])
    eng = Engine(U.default_sheet_config(), w, h, U.FMT_GRAY8, group_pages=2, lanes=2)
    out1, res1 = eng.process_numpy(pages)
    out2, res2 = eng.process_numpy(pages)
    assert np.array_equal(out1, out2)
    out3, res3 = eng.process_numpy(out1)
    eng.close()
    for r in res1:
        assert r.status == 0 and r.deskew_mask_count == 1 and r.rotation[0] != 0.0
    for r in res3:
        assert abs(r.rotation[0]) <= np.deg2rad(0.2) + 1e-6


# ---- BASELINE.json configs 3 and 4 at full size -------------------------------------
# The reference needs minutes per sheet here, so full-size runs are held to
# size-independent properties; the same configurations are compared bit for bit
# with the reference at reduced size above (test_engine_rgb_small, _double_layout).

def test_engine_color_a4_full_size(ref_lib):
    """Config 3: A4 RGB24, grayfilter + blurfilter + cubic deskew — one sheet against
    the reference (about 40 s of CPU), then determinism on a second run."""
    from unpaper_gpu_b200.lib import Engine
    w, h = synth.A4_W, synth.A4_H
    pages = synth.color_page(1, w, h)[None]
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = 1
    out, res = _compare(cfg, pages, w, h, U.FMT_RGB24, ref_lib, group=1, lanes=1)
    eng = Engine(cfg, w, h, U.FMT_RGB24, group_pages=1, lanes=1)
    out2, _ = eng.process_numpy(pages)
    eng.close()
    assert np.array_equal(out, out2)


def test_engine_double_600dpi_properties():
    """Config 4: 7016x4960 GRAY8 two-page sheet, --layout double."""
    from unpaper_gpu_b200.lib import Engine
    w, h = 7016, 4960
    pages = np.stack([synth.double_sheet(i, w, h) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.layout = U.LAYOUT_DOUBLE
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=2, lanes=1)
    out, res = eng.process_numpy(pages)
    out2, res2 = eng.process_numpy(pages)
    eng.close()
    assert np.array_equal(out, out2)
    for r in res:
        assert r.status == 0
        assert r.deskew_mask_count == 2 and r.center_mask_count == 2 and r.border_count == 2
        for k in range(2):
            x0, y0, x1, y1 = U.rect_tuple(r.deskew_masks[k])
            assert 0 <= x0 < x1 < w and y0 == 0 and y1 == h - 1
            assert abs(r.rotation[k]) <= np.deg2rad(5.1) + 1e-6
        # the two page masks do not overlap and sit in their own halves
        assert U.rect_tuple(r.deskew_masks[0])[2] < w // 2 + 60 < U.rect_tuple(r.deskew_masks[1])[2]
    # dark scan edges are gone, the sheet is mostly white, ink survives
    assert (out[:, :, :30] == 255).all()
    frac_dark = (out < 128).mean()
    assert 0.01 < frac_dark < 0.2


def test_engine_pre_post_mirror_shift(ref_lib):
    """Size-preserving geometry options of the pre and post stages
    (sheet_stages.c:200-208, :499-508): mirror, then shift_image."""
    w, h = 620, 877
    pages = np.stack([synth.gray_page(90 + i, w, h, box=SMALL_BOX) for i in range(3)])
    for pm, ps, qm, qs in (((True, False), (7, -5), (False, True), (-3, 11)),
                           ((True, True), (0, 0), (False, False), (12, 0)),
                           ((False, False), (-9, 4), (True, True), (0, 0))):
        cfg = U.default_sheet_config()
        cfg.pre_mirror, cfg.pre_shift = U.Direction(*pm), U.Delta(*ps)
        cfg.post_mirror, cfg.post_shift = U.Direction(*qm), U.Delta(*qs)
        _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)
    cfg = U.default_sheet_config()
    cfg.no_blackfilter = cfg.no_noisefilter = 1
    cfg.pre_mirror, cfg.post_shift = U.Direction(False, True), U.Delta(5, -7)
    rgb = np.stack([synth.color_page(i, w, h) for i in range(2)])
    _compare(cfg, rgb, w, h, U.FMT_RGB24, ref_lib, group=2, lanes=1)


def test_engine_two_pages_per_sheet(ref_lib):
    """input_count = 2: two decoded pages are placed side by side on one sheet
    (sheet_stages.c:140-165), double layout — through host buffers and device-resident."""
    import torch
    from unpaper_gpu_b200.lib import Engine
    w, h = 620, 877
    pages = np.stack([synth.gray_page(120 + i, w, h, box=SMALL_BOX) for i in range(6)])   # 3 sheets x 2 pages
    cfg = U.default_sheet_config()
    cfg.input_count, cfg.layout = 2, U.LAYOUT_DOUBLE
    out, res = _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)
    assert out.shape == (3, h, 2 * w)
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=2, lanes=1)
    d_in = torch.from_numpy(pages.reshape(-1)).cuda()
    d_out = torch.empty(out.size, dtype=torch.uint8, device="cuda")
    eng.process_ptr(d_in.data_ptr(), d_out.data_ptr(), 3, False, None)
    eng.close()
    assert np.array_equal(d_out.cpu().numpy().reshape(out.shape), out)


def _wipes_borders_cfg():
    cfg = U.default_sheet_config()
    cfg.pre_wipe_count = 1; cfg.pre_wipes[0] = U.rect(100, 120, 160, 170)
    cfg.wipe_count = 2; cfg.wipes[0] = U.rect(300, 400, 340, 460); cfg.wipes[1] = U.rect(-5, 800, 50, 900)
    cfg.post_wipe_count = 1; cfg.post_wipes[0] = U.rect(500, 50, 619, 90)
    cfg.pre_border = U.Border(3, 4, 5, 6); cfg.border = U.Border(10, 0, 0, 12); cfg.post_border = U.Border(0, 7, 8, 0)
    cfg.pre_mask_count = 1; cfg.pre_masks[0] = U.rect(20, 20, 600, 860)
    cfg.noisefilter_intensity = 6
    return cfg


def test_engine_wipes_borders_premasks_points(ref_lib):
    """The static rectangles of the pre/mid/post stages (pre-masks, wipes — one of them
    partly outside the sheet —, borders), a non-default noise intensity, then a user
    point, a gray mask colour and an off-white sheet background."""
    w, h = 620, 877
    pages = np.stack([synth.gray_page(140 + i, w, h, box=SMALL_BOX) for i in range(3)])
    _compare(_wipes_borders_cfg(), pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)
    cfg = _wipes_borders_cfg()
    cfg.point_count = 1; cfg.points[0] = U.Point(300, 430)
    cfg.mask_color = U.Pixel(200, 200, 200); cfg.sheet_background = U.Pixel(250, 250, 250)
    _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=3, lanes=1)


def test_engine_reference_golden_c1():
    """The reference's own exact golden on this path (tests/unpaper_tests.py:568-599:
    mask + border scan in both directions, pre-wipe, pre-border, RGB24) through the CUDA
    engine: every byte equals goldenC1.ppm (fixture tests/golden/c1_fixture.npz)."""
    import os
    import golden_cases as G
    from unpaper_gpu_b200.lib import Engine
    f = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_fixture.npz"))
    page, golden = f["page"], f["golden"]
    h, w, _ = page.shape
    pages = np.stack([page.reshape(h, 3 * w)] * 3)
    eng = Engine(G.c1_config(), w, h, U.FMT_RGB24, group_pages=2, lanes=2)
    out, res = eng.process_numpy(pages)
    eng.close()
    for i in range(3):
        assert res[i].status == 0
        assert np.array_equal(out[i].reshape(h, w, 3), golden), f"sheet {i}"
        assert U.border_tuple(res[i].borders[0]) == (35, 20, 16, 21)


def _overlap_cfg(edges=(False, False, True, False)):
    """Two user points whose masks hit the maximum-width fallback (masks.c:149-169) and
    therefore overlap: detect_rotation(mask 1) must see mask 0 already deskewed
    (sheet_stages.c:406-413).  One scan edge, because the other edge of each mask cuts
    through the text."""
    cfg = U.default_sheet_config()
    cfg.point_count = 2
    cfg.points[0] = U.Point(250, 430); cfg.points[1] = U.Point(370, 430)
    cfg.mask_detection.maximum_width = 300
    cfg.deskew.scan_edges = U.Edges(*edges)
    return cfg


def test_engine_overlapping_masks(ref_lib):
    w, h = 620, 877
    pages = np.stack([synth.gray_page(200 + i, w, h, box=SMALL_BOX) for i in range(4)])
    for edges in ((False, False, True, False), (True, False, False, False)):
        out, res = _compare(_overlap_cfg(edges), pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)
        overl = 0
        for r in res:
            assert r.deskew_mask_count == 2
            a, b = U.rect_tuple(r.deskew_masks[0]), U.rect_tuple(r.deskew_masks[1])
            overl += a != b and a[2] >= b[0] and r.rotation[0] != 0.0
        assert overl >= 3, "the fixture no longer produces overlapping, rotated masks"
    # double layout on a width divisible by 4: the fallback masks share column W/2
    w2 = 1240
    pages2 = np.stack([synth.double_sheet(210 + i, w2, h) for i in range(2)])
    cfg = U.default_sheet_config()
    cfg.layout = U.LAYOUT_DOUBLE
    cfg.mask_detection.maximum_width = 400
    _compare(cfg, pages2, w2, h, U.FMT_GRAY8, ref_lib, group=2, lanes=1)


@pytest.mark.parametrize("edges", [(True, True, True, True), (True, True, True, False), (False, True, False, True)])
def test_engine_deskew_scan_edges(ref_lib, edges):
    """3-4 scan edges: average / deviation / sinf / cosf of deskew.c:218-261 come from the
    host's libm through a stream-ordered host function; top/bottom edges use the sampling kernel."""
    w, h = 620, 877
    pages = np.stack([synth.gray_page(220 + i, w, h, box=SMALL_BOX) for i in range(3)])
    cfg = U.default_sheet_config()
    cfg.deskew.scan_edges = U.Edges(*edges)
    out, res = _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=2, lanes=2)


def test_engine_per_sheet_switch_lists(ref_lib):
    """options->no_*_multi_index with sheet lists + ignore_multi_index (isExcluded(),
    sheet_stages.c:282-493, :644-650): stages are skipped for the listed sheet numbers only."""
    w, h = 620, 877
    pages = np.stack([synth.gray_page(240 + i, w, h, box=SMALL_BOX) for i in range(7)])
    cfg = U.default_sheet_config()
    cfg.first_sheet_nr = 3                                   # sheets 3..9
    cfg.no_deskew_sheets = U.multi_index([4, 8])
    cfg.no_noisefilter_sheets = U.multi_index([3, 4])
    cfg.no_mask_center_sheets = U.multi_index([6])
    cfg.no_border_align_sheets = U.multi_index([6, 7])
    cfg.no_blackfilter_sheets = U.multi_index([9])
    cfg.ignore_sheets = U.multi_index([5])
    cfg.wipe_count = 1; cfg.wipes[0] = U.rect(300, 400, 340, 460)
    cfg.no_wipe_sheets = U.multi_index([7])
    cfg.border = U.Border(10, 0, 0, 12)
    cfg.no_border_sheets = U.multi_index([8])
    out, res = _compare(cfg, pages, w, h, U.FMT_GRAY8, ref_lib, group=4, lanes=2)
    assert res[1].deskew_mask_count == 0 and res[0].deskew_mask_count == 1
    assert res[2].deskew_mask_count == 0 and res[2].border_count == 0      # sheet 5: everything off
    cfg2 = U.default_sheet_config()
    cfg2.no_grayfilter_sheets = U.multi_index(None)          # count -1: every sheet
    cfg3 = U.default_sheet_config()
    cfg3.no_grayfilter = 1
    from unpaper_gpu_b200.lib import Engine
    outs = []
    for c in (cfg2, cfg3):
        eng = Engine(c, w, h, U.FMT_GRAY8, group_pages=4, lanes=1)
        outs.append(eng.process_numpy(pages[:2])[0])
        eng.close()
    assert np.array_equal(outs[0], outs[1])


def test_engine_rejects_bad_counts():
    """Counts outside the fixed arrays of B200SheetConfig are refused at creation."""
    from unpaper_gpu_b200.lib import Engine
    for field, v in (("point_count", 9), ("pre_mask_count", -1), ("wipe_count", 100), ("post_wipe_count", 9),
                     ("pre_wipe_count", -3), ("output_count", 3), ("input_count", 3)):
        cfg = U.default_sheet_config()
        setattr(cfg, field, v)
        with pytest.raises(RuntimeError, match="engine"):
            Engine(cfg, 64, 64, U.FMT_GRAY8, group_pages=1, lanes=1)


# ---- the reference's own real scans through the CUDA engine --------------------------
# tests/golden/*.npz hold the reference's 1-bit test scans (packed); golden_vectors.json
# holds what the UNMODIFIED reference (oracle/_ref) produced for them when the vectors
# were made (tests/golden/make_golden.py), incl. its distance to the reference's golden images.

def _golden(name):
    import json
    import os
    d = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(d, "golden_vectors.json")) as f:
        vec = json.load(f)
    return vec, np.load(os.path.join(d, name))


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_engine_reference_golden_a1():
    """BASELINE config 1 = the reference's test_a1 (tests/unpaper_tests.py:653-669): default
    pipeline on imgsrc001.  Every byte of the sheet and every decision equal the reference
    CPU backend's; the sheet is within 1e-4 of goldenA1.pbm at unpaper's own pbm threshold."""
    import golden_cases as G
    from unpaper_gpu_b200.lib import Engine
    vec, f = _golden("a1_fixture.npz")
    w, h = (int(v) for v in f["size"])
    page = (np.unpackbits(f["page_bits"], axis=1)[:, :w] * 255).astype(np.uint8)
    assert _sha(page) == vec["A1"]["input_sha256"]
    cfg = U.default_sheet_config()
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=1, lanes=1)
    out, res = eng.process_numpy(page[None])
    eng.close()
    assert res[0].status == 0
    assert G.result_dict(res[0]) == vec["A1"]["result"]
    assert _sha(out[0]) == vec["A1"]["output_sha256"]
    gold_black = np.unpackbits(f["golden_bits"], axis=1)[:, :w] == 1
    ratio = float(np.mean((out[0] < cfg.abs_black_threshold) != gold_black))
    assert ratio < 1e-4 and abs(ratio - vec["A1"]["golden_diff_ratio_thr170"]) < 1e-12
    # the same scan as a 1-bit page, written as pbm (the reference test's actual file type)
    eng = Engine(cfg, w, h, U.FMT_MONOBLACK, group_pages=1, lanes=1)
    mono, res = eng.process_numpy(f["page_bits"][None])
    eng.close()
    assert float(np.mean((np.unpackbits(mono[0], axis=1)[:, :w] == 1) != gold_black)) < 1e-4


def test_engine_reference_golden_f3():
    """The reference's test_f3 (:787-810): two 1-bit scans merged on one double-layout sheet
    (--input-pages 2); the reference build reproduces goldenF.pbm exactly, and so must this."""
    import golden_cases as G
    from unpaper_gpu_b200.lib import Engine
    vec, f = _golden("e_fixture.npz")
    w, h = (int(v) for v in f["size"])
    bits = f["pages_bits"]
    assert [_sha(b) for b in bits[:2]] == vec["F3"]["input_sha256"]
    cfg = U.default_sheet_config()
    cfg.layout, cfg.input_count = U.LAYOUT_DOUBLE, 2
    eng = Engine(cfg, w, h, U.FMT_MONOBLACK, group_pages=1, lanes=1)
    out, res = eng.process_numpy(bits[:2].reshape(1, -1))
    eng.close()
    assert res[0].status == 0 and G.result_dict(res[0]) == vec["F3"]["result"]
    assert vec["F3"]["differing_pixels_vs_goldenF"] == 0
    assert _sha(out[0]) == vec["F3"]["monowhite_sha256"]


def test_engine_reference_golden_e1():
    """The reference's test_e1 (:763-783): --layout double --output-pages 2 on three 1-bit
    double-page scans; the six pbm files equal the ones the reference's own output stage wrote
    (sheet split + saveImage), which are within 1e-4 of goldenE1-0N.pbm."""
    import golden_cases as G
    from unpaper_gpu_b200.lib import Engine
    vec, f = _golden("e_fixture.npz")
    w, h = (int(v) for v in f["size"])
    cfg = U.default_sheet_config()
    cfg.layout, cfg.output_count = U.LAYOUT_DOUBLE, 2
    eng = Engine(cfg, w, h, U.FMT_MONOBLACK, group_pages=2, lanes=2)
    assert [eng.out_w, eng.sheet_h] == vec["E1"]["split_size"] and eng.out_fmt == U.FMT_MONOWHITE
    out, res = eng.process_numpy(f["pages_bits"].reshape(3, -1))
    eng.close()
    assert max(vec["E1"]["split_files_golden_diff_ratio"]) < 1e-4
    for k in range(3):
        assert res[k].status == 0 and G.result_dict(res[k]) == vec["E1"]["results"][k]
        for half in range(2):
            assert _sha(out[k, half]) == vec["E1"]["split_files_sha256"][2 * k + half], f"sheet {k} page {half}"


@pytest.mark.timeout(900)
def test_engine_double_600dpi_vs_reference(ref_lib):
    """BASELINE config 4 at full size (7016x4960, --layout double, sheet split): one sheet bit
    for bit against the reference's process_sheet() + output stage (about a minute of CPU)."""
    from oracle import checker
    from unpaper_gpu_b200.lib import Engine
    w, h = 7016, 4960
    pages = synth.double_sheet(1, w, h)[None]
    cfg = U.default_sheet_config()
    cfg.layout, cfg.output_count = U.LAYOUT_DOUBLE, 2
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=1, lanes=1)
    out, res = eng.process_numpy(pages)
    eng.close()
    files, rres = checker.process_sheets_files_cpu(ref_lib, cfg, pages, w, h, U.FMT_GRAY8, output_count=2)
    a, b = res[0], rres[0]
    assert a.status == 0 and b.status == 0 and a.deskew_mask_count == b.deskew_mask_count == 2
    for k in range(2):
        assert U.rect_tuple(a.deskew_masks[k]) == U.rect_tuple(b.deskew_masks[k])
        assert a.rotation[k] == b.rotation[k] and a.rotation[k] != 0.0
        assert U.rect_tuple(a.center_masks[k]) == U.rect_tuple(b.center_masks[k])
        assert U.border_tuple(a.borders[k]) == U.border_tuple(b.borders[k])
        assert files[0][k][:3] == (U.FMT_GRAY8, w // 2, h)
        assert np.array_equal(out[0, k], files[0][k][3]), f"page {k}: {int((out[0, k] != files[0][k][3]).sum())} differing bytes"


def _size_cases():
    """name -> (config edits, page format, input_count): the size-changing options of the decode, pre and
    post stages (sheet_stages.c:134-145, :216-230, :511-531)."""
    RS = U.RectangleSize
    return {
        "pre_rotate_cw": (dict(pre_rotate=90), U.FMT_GRAY8, 1),
        "pre_rotate_ccw_two_pages": (dict(pre_rotate=-90, layout=U.LAYOUT_DOUBLE), U.FMT_GRAY8, 2),
        "post_rotate": (dict(post_rotate=-90), U.FMT_GRAY8, 1),
        "stretch": (dict(stretch_size=RS(700, 940)), U.FMT_GRAY8, 1),
        "stretch_width_only_linear": (dict(stretch_size=RS(560, -1), interpolate_type=U.INTERP_LINEAR), U.FMT_GRAY8, 1),
        "pre_zoom": (dict(pre_zoom_factor=1.25), U.FMT_GRAY8, 1),
        "post_zoom_and_post_stretch": (dict(post_zoom_factor=0.5, post_stretch_size=RS(900, 1000)), U.FMT_GRAY8, 1),
        "page_size_wider": (dict(page_size=RS(900, 877)), U.FMT_GRAY8, 1),
        "page_size_smaller": (dict(page_size=RS(500, 800)), U.FMT_GRAY8, 1),
        "post_page_size": (dict(post_page_size=RS(640, 640)), U.FMT_GRAY8, 1),
        "sheet_size_larger": (dict(sheet_size=RS(700, 1000)), U.FMT_GRAY8, 1),
        "sheet_size_crops": (dict(sheet_size=RS(600, 850)), U.FMT_GRAY8, 1),
        "everything_rgb": (dict(pre_rotate=90, stretch_size=RS(940, 700), post_rotate=90, post_page_size=RS(720, 960),
                                no_blackfilter=1, no_noisefilter=1), U.FMT_RGB24, 1),
        "mono_pages_rotated": (dict(pre_rotate=90, post_rotate=-90), U.FMT_MONOBLACK, 1),
    }


def _size_case(name):
    edits, fmt, ic = _size_cases()[name]
    cfg = U.default_sheet_config()
    cfg.input_count = ic
    for k, v in edits.items():
        setattr(cfg, k, v)
    w, h = (620, 877) if fmt != U.FMT_MONOBLACK else (624, 880)
    n = 2 * ic
    if fmt == U.FMT_RGB24:
        pages = np.stack([synth.color_page(600 + i, w, h) for i in range(n)])
    else:
        pages = np.stack([synth.gray_page(600 + i, w, h, box=SMALL_BOX) for i in range(n)])
        if fmt == U.FMT_MONOBLACK:
            pages = np.stack([np.packbits(p >= 128, axis=1) for p in pages])
    return cfg, pages, w, h, fmt


@pytest.mark.parametrize("name", sorted(_size_cases()))
def test_engine_size_changing_options(ref_lib, name):
    cfg, pages, w, h, fmt = _size_case(name)
    if fmt == U.FMT_MONOBLACK:
        from unpaper_gpu_b200.lib import Engine
        eng = Engine(cfg, w, h, fmt, group_pages=2, lanes=1)
        out, res = eng.process_numpy(pages)
        size = (eng.sheet_w, eng.sheet_h)
        eng.close()
        rout, rres = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, fmt, threads=4, out_size=size)
        assert all(r.status == 0 for r in res) and all(r.status == 0 for r in rres)
        assert size[0] % 8 == 0
        assert np.array_equal(out, rout ^ 0xFF)      # the harness hands back MONOBLACK; saveImage() writes MONOWHITE
        return
    out, res = _compare(cfg, pages, w, h, fmt, ref_lib, group=2, lanes=2)
    assert (res[0].sheet_width, res[0].sheet_height) == (out.shape[2] // (3 if fmt == U.FMT_RGB24 else 1), out.shape[1])
