"""Shared helpers for the parity tests: seeded images and exact comparison."""
import ctypes as C
import hashlib

import numpy as np

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth

FMTS_BYTE = [U.FMT_GRAY8, U.FMT_RGB24, U.FMT_Y400A]
FMTS_ALL = FMTS_BYTE + [U.FMT_MONOWHITE, U.FMT_MONOBLACK]
FMT_NAME = {U.FMT_GRAY8: "gray8", U.FMT_RGB24: "rgb24", U.FMT_Y400A: "y400a",
            U.FMT_MONOWHITE: "monowhite", U.FMT_MONOBLACK: "monoblack"}


def linesize(fmt, w):
    return (U.bytes_per_row(fmt, w) + 7) // 8 * 8


def noise_image(seed, w, h, fmt, dark=0.05, lo=0, hi=256, bg=255):
    """uint8 [h, linesize]: background `bg`, a fraction `dark` of pixels random in [lo,hi)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ls = linesize(fmt, w)
    img = np.zeros((h, ls), dtype=np.uint8)
    m = rng.random((h, w)) < dark
    if fmt in (U.FMT_MONOWHITE, U.FMT_MONOBLACK):
        bits = m if fmt == U.FMT_MONOWHITE else ~m      # monowhite: set bit = black
        packed = np.packbits(bits, axis=1)
        img[:, :packed.shape[1]] = packed
        return img
    bpp = {U.FMT_GRAY8: 1, U.FMT_Y400A: 2, U.FMT_RGB24: 3}[fmt]
    v = np.full((h, w, bpp), bg, dtype=np.uint8)
    vals = rng.integers(lo, hi, size=(h, w, bpp), dtype=np.uint8)
    if fmt == U.FMT_RGB24:
        v[m] = vals[m]
    else:
        v[m, 0] = vals[m, 0]
        if fmt == U.FMT_Y400A:
            v[..., 1] = rng.integers(0, 256, size=(h, w), dtype=np.uint8)  # arbitrary alpha
    img[:, :w * bpp] = v.reshape(h, w * bpp)
    return img


def blobs_image(seed, w, h, fmt, nblobs=6, texture=True):
    """Black bands/blobs on light, slightly textured paper (blackfilter food)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = np.full((h, w), 255, dtype=np.uint8)
    if texture:
        t = rng.random((h, w)) < 0.03
        g[t] = rng.integers(175, 250, size=int(t.sum()), dtype=np.uint8)
    for _ in range(nblobs):
        bw, bh = int(rng.integers(5, max(6, w // 3))), int(rng.integers(5, max(6, h // 3)))
        x, y = int(rng.integers(-bw // 2, w - bw // 2)), int(rng.integers(-bh // 2, h - bh // 2))
        g[max(y, 0):y + bh, max(x, 0):x + bw] = rng.integers(0, 40)
    sp = rng.random((h, w)) < 0.002
    g[sp] = 0
    return from_gray(g, fmt)


def from_gray(g, fmt):
    h, w = g.shape
    ls = linesize(fmt, w)
    img = np.zeros((h, ls), dtype=np.uint8)
    if fmt == U.FMT_GRAY8:
        img[:, :w] = g
    elif fmt == U.FMT_RGB24:
        img[:, :3 * w] = np.repeat(g, 3, axis=1)
    elif fmt == U.FMT_Y400A:
        v = np.stack([g, np.full_like(g, 255)], axis=-1)
        img[:, :2 * w] = v.reshape(h, 2 * w)
    else:
        black = g < 128
        bits = black if fmt == U.FMT_MONOWHITE else ~black
        p = np.packbits(bits, axis=1)
        img[:, :p.shape[1]] = p
    return img


def visible(img, fmt, w):
    """The bytes (and bits) that carry pixels: ignores row padding."""
    n = U.bytes_per_row(fmt, w)
    v = img[:, :n].copy()
    if fmt in (U.FMT_MONOWHITE, U.FMT_MONOBLACK) and w % 8:
        v[:, -1] &= (0xFF << (8 - w % 8)) & 0xFF
    return v


def assert_same(a, b, fmt, w, what):
    va, vb = visible(a, fmt, w), visible(b, fmt, w)
    if np.array_equal(va, vb):
        return
    d = np.argwhere(va != vb)
    msg = [f"{what}: {len(d)} differing bytes of {va.size} ({FMT_NAME[fmt]} {w}x{a.shape[0]})"]
    for y, x in d[:12]:
        msg.append(f"  byte ({x},{y}): cuda={va[y, x]} ref={vb[y, x]}")
    raise AssertionError("\n".join(msg))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def himg(a, fmt, w, bg=(255, 255, 255), abt=170):
    return U.HostOps.himg(a, fmt, w, bg, abt)


def run_inplace(ops, name, img, fmt, w, *args, abt=170, bg=(255, 255, 255)):
    out = img.copy()
    hi = himg(out, fmt, w, bg, abt)
    ops.call(name, C.byref(hi), *args)
    return out
