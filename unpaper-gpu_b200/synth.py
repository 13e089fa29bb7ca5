"""Seeded synthetic scanned pages of the shapes BASELINE.json names.

Every page is a pure function of (kind, index, width, height): a white sheet,
a text-like content box rotated by a per-page skew in [-5 deg, +5 deg],
speckle noise, and dark scan edges separated from the content by a white
gutter (the reference's detect_edge() only terminates on a light bar inside
the image — SURVEY.md section 0 item 5).

The generator is numpy (vectorised); the per-page seed is
0x9E3779B9 * (index + 1) mod 2**32 fed to PCG64.
"""
import numpy as np

A4_W, A4_H = 2480, 3508


def _rng(index, salt=0):
    return np.random.Generator(np.random.PCG64(((0x9E3779B9 * (index + 1)) ^ salt) & 0xFFFFFFFF))


def _text_block(rng, w, h, scale, color=False):
    """Unrotated content box: glyph grid, returns uint8 [h, w] or [h, w, 3]."""
    line_pitch = max(8, int(round(50 * scale)))
    x_height = max(4, int(round(18 * scale)))
    glyph_pitch = max(5, int(round(23 * scale)))
    glyph_w = max(3, int(round(15 * scale)))
    nlines = max(1, h // line_pitch)
    nglyph = max(1, w // glyph_pitch)
    bx, by = max(1, glyph_w // 5), max(1, x_height // 6)
    # 5x6 block pattern per glyph, each block bx*by pixels; keep ~60 % of blocks
    pat = rng.random((nlines, nglyph, 6, 5)) < 0.6
    pat[:, :, :, 0] |= rng.random((nlines, nglyph, 1)) < 0.7  # a stem: big components
    present = rng.random((nlines, nglyph)) >= 0.125           # 1/8 dropout (spaces)
    pat &= present[:, :, None, None]
    if color:
        ink = rng.integers(0, 91, size=(nlines, nglyph, 3), dtype=np.uint8)
        img = np.full((h, w, 3), 255, dtype=np.uint8)
    else:
        ink = rng.integers(20, 52, size=(nlines, nglyph), dtype=np.uint8)
        img = np.full((h, w), 255, dtype=np.uint8)
    cell = np.kron(pat, np.ones((by, bx), dtype=bool))  # [L, G, 6*by, 5*bx]
    gh, gw = cell.shape[2], cell.shape[3]
    for li in range(nlines):
        y0 = li * line_pitch + (line_pitch - gh) // 2
        if y0 + gh > h:
            break
        for gi in range(nglyph):
            x0 = gi * glyph_pitch
            if x0 + gw > w or not present[li, gi]:
                continue
            m = cell[li, gi]
            sub = img[y0:y0 + gh, x0:x0 + gw]
            sub[m] = ink[li, gi]
    return img


def _rotate_into(page, block, theta, cx, cy):
    """Nearest-neighbour place `block`, rotated by theta about (cx, cy), on page."""
    h, w = page.shape[:2]
    bh, bw = block.shape[:2]
    half_w, half_h = bw / 2.0, bh / 2.0
    r = int(np.ceil(np.hypot(half_w, half_h))) + 2
    x0, x1 = max(0, int(cx) - r), min(w, int(cx) + r)
    y0, y1 = max(0, int(cy) - r), min(h, int(cy) + r)
    ys, xs = np.mgrid[y0:y1, x0:x1].astype(np.float32)
    c, s = np.float32(np.cos(theta)), np.float32(np.sin(theta))
    dx, dy = xs - np.float32(cx), ys - np.float32(cy)
    sx = np.rint(dx * c + dy * s + np.float32(half_w)).astype(np.int32)
    sy = np.rint(-dx * s + dy * c + np.float32(half_h)).astype(np.int32)
    ok = (sx >= 0) & (sx < bw) & (sy >= 0) & (sy < bh)
    sub = page[y0:y1, x0:x1]
    sub[ok] = block[sy[ok], sx[ok]]


def gray_page(index, width=A4_W, height=A4_H, max_skew_deg=5.0, speckle=1.0 / 5000,
              dark_edges=True, box=(0.76, 0.80)):
    """BASELINE config 2 ("C2"): GRAY8 page, text box +-5 deg, speckle, dark edges."""
    rng = _rng(index)
    scale = width / float(A4_W)
    page = np.full((height, width), 255, dtype=np.uint8)
    bw, bh = int(width * box[0]), int(height * box[1])
    block = _text_block(rng, bw, bh, scale)
    theta = np.deg2rad(rng.uniform(-max_skew_deg, max_skew_deg))
    _rotate_into(page, block, theta, width / 2.0, height / 2.0)
    if speckle > 0:
        n = rng.binomial(width * height, speckle)
        page[rng.integers(0, height, n), rng.integers(0, width, n)] = 0
    if dark_edges:
        # never narrower than the blackfilter's 20-px scan bar, or the filter
        # leaves them in and detect_edge() runs off the sheet (it never returns
        # in the reference: masks.c:88-97)
        le, re = max(24, int(round(40 * scale))), max(22, int(round(30 * scale)))
        page[:, :le] = 10
        page[:, width - re:] = 10
    return page


def color_page(index, width=A4_W, height=A4_H, max_skew_deg=5.0):
    """BASELINE config 3 ("C3"): RGB24 page with dark coloured glyphs, light-gray
    blotches (grayfilter food) and sparse tinted blocks (blurfilter food)."""
    rng = _rng(index, salt=0xC3C3C3)
    scale = width / float(A4_W)
    page = np.full((height, width, 3), 255, dtype=np.uint8)
    for _ in range(12):
        bw_, bh_ = (int(rng.integers(60, 201) * scale) + 1 for _ in range(2))
        x, y = int(rng.integers(0, max(1, width - bw_))), int(rng.integers(0, max(1, height - bh_)))
        page[y:y + bh_, x:x + bw_] = rng.integers(180, 231)
    bw, bh = int(width * 0.76), int(height * 0.80)
    block = _text_block(rng, bw, bh, scale, color=True)
    theta = np.deg2rad(rng.uniform(-max_skew_deg, max_skew_deg))
    _rotate_into(page, block, theta, width / 2.0, height / 2.0)
    for _ in range(40):
        x, y = int(rng.integers(0, width - 4)), int(rng.integers(0, height - 4))
        page[y:y + 3, x:x + 3] = rng.integers(100, 200, size=3)
    return page.reshape(height, width * 3)


def double_sheet(index, width=7016, height=4960, max_skew_deg=5.0, speckle=1.0 / 5000):
    """BASELINE config 4 ("C4"): one GRAY8 scan holding two pages side by side,
    each with its own skew."""
    rng = _rng(index, salt=0xD0B1E)
    scale = (width / 2.0) / float(A4_W) * 0.7
    sheet = np.full((height, width), 255, dtype=np.uint8)
    for k in range(2):
        bw, bh = int(width * 0.5 * 0.70), int(height * 0.78)
        block = _text_block(rng, bw, bh, scale)
        theta = np.deg2rad(rng.uniform(-max_skew_deg, max_skew_deg))
        _rotate_into(sheet, block, theta, width * (0.25 + 0.5 * k), height / 2.0)
    if speckle > 0:
        n = rng.binomial(width * height, speckle)
        sheet[rng.integers(0, height, n), rng.integers(0, width, n)] = 0
    le = max(2, int(round(40 * scale)))
    sheet[:, :le] = 10
    sheet[:, width - le:] = 10
    return sheet


def random_image(seed, width, height, fmt_bpp=1, dark_frac=0.02, levels=(0, 256)):
    """Unstructured noise image for op-level property tests."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = np.full((height, width * fmt_bpp), 255, dtype=np.uint8)
    m = rng.random((height, width)) < dark_frac
    vals = rng.integers(levels[0], levels[1], size=(height, width, fmt_bpp), dtype=np.uint8)
    v = img.reshape(height, width, fmt_bpp)
    v[m] = vals[m]
    return img
