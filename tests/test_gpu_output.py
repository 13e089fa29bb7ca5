"""Output side: sheet_stage_output's pixel-format conversion on the
device and the direct PNM writer — reference file.c:134-260.

The checker is the reference's own saveImage() (file.c compiled unmodified into
oracle/_ref against oracle/shim_codec; saveImageDirect writes the PNM the test reads
back) — `checker.save_image_cpu`, `checker.process_sheets_files_cpu`."""
import ctypes as C
import os

import numpy as np
import pytest

import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from util import himg, linesize, noise_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc_ops():
    from oracle import checker
    lib = checker.load_oracle()
    if lib is None:
        pytest.skip("oracle/liboracle.so not built")
    return U.HostOps(lib, "orc_host_")


@pytest.mark.parametrize("sfmt,dfmt", [(U.FMT_GRAY8, U.FMT_MONOWHITE), (U.FMT_RGB24, U.FMT_MONOWHITE),
                                       (U.FMT_MONOBLACK, U.FMT_MONOWHITE), (U.FMT_Y400A, U.FMT_MONOWHITE),
                                       (U.FMT_RGB24, U.FMT_GRAY8), (U.FMT_GRAY8, U.FMT_RGB24),
                                       (U.FMT_MONOWHITE, U.FMT_GRAY8), (U.FMT_MONOBLACK, U.FMT_RGB24),
                                       (U.FMT_Y400A, U.FMT_GRAY8), (U.FMT_GRAY8, U.FMT_GRAY8),
                                       (U.FMT_MONOWHITE, U.FMT_MONOWHITE)])
@pytest.mark.parametrize("w,h", [(37, 29), (203, 77), (640, 64)])
@pytest.mark.parametrize("abt", [170, 84])
def test_convert_format(cuda_ops, ref_lib, sfmt, dfmt, w, h, abt):
    from oracle import checker
    src = noise_image(11, w, h, sfmt, dark=0.4)
    d = np.full((h, linesize(dfmt, w)), 0x5A, dtype=np.uint8)
    cuda_ops.call("convert_format", C.byref(himg(src, sfmt, w, abt=abt)), C.byref(himg(d, dfmt, w, abt=abt)))
    rfmt, rw, rh, ref = checker.save_image_cpu(ref_lib, himg(src, sfmt, w, abt=abt), dfmt)   # file.c:186-262
    assert (rfmt, rw, rh) == (dfmt, w, h)
    outs = [d]
    row = U.bytes_per_row(dfmt, w)
    a, b = d[:, :row].copy(), ref.copy()
    if dfmt == U.FMT_MONOWHITE and sfmt not in (U.FMT_GRAY8, U.FMT_RGB24, U.FMT_MONOBLACK) and w % 8:
        # generic branch (file.c:257-259): set_pixel only touches the w real pixels of a
        # create_image(fill=false) buffer, the tail bits of the last byte are unspecified
        keep = (0xFF << (8 - w % 8)) & 0xFF
        a[:, -1] &= keep
        b[:, -1] &= keep
    assert np.array_equal(a, b)
    assert (outs[0][:, row:] == 0x5A).all()     # nothing written past the row


def test_output_format_mapping_and_header():
    from unpaper_gpu_b200 import lib as L
    lib = L.load()
    assert lib.unpaper_b200_output_format(U.FMT_Y400A) == U.FMT_GRAY8          # file.c:201-208
    assert lib.unpaper_b200_output_format(U.FMT_MONOBLACK) == U.FMT_MONOWHITE
    assert lib.unpaper_b200_output_format(U.FMT_RGB24) == U.FMT_RGB24
    buf = C.create_string_buffer(64)
    for fmt, want in ((U.FMT_GRAY8, b"P5\n31 7\n255\n"), (U.FMT_RGB24, b"P6\n31 7\n255\n"), (U.FMT_MONOWHITE, b"P4\n31 7\n")):
        n = lib.unpaper_b200_pnm_header(fmt, 31, 7, buf, 64)
        assert buf.raw[:n] == want
    assert lib.unpaper_b200_pnm_header(U.FMT_Y400A, 31, 7, buf, 64) < 0


def test_engine_mono_output(ref_lib, tmp_path):
    """pbm output: the engine converts on the device, D2H carries 1 bit/px; the bytes
    equal what the reference's process_sheet() + saveImage() write for the same pages
    with a MONOBLACK (-> MONOWHITE, file.c:205-207) output format."""
    from oracle import checker
    from unpaper_gpu_b200.lib import Engine
    from unpaper_gpu_b200 import lib as L
    w, h = 620, 877
    pages = np.stack([synth.gray_page(i, w, h, box=(0.60, 0.72)) for i in range(5)])
    cfg = U.default_sheet_config()
    eng = Engine(cfg, w, h, U.FMT_GRAY8, group_pages=2, lanes=2)
    gray, _ = eng.process_numpy(pages)
    eng.set_output_format(U.FMT_MONOBLACK)      # written as MONOWHITE
    assert eng.out_fmt == U.FMT_MONOWHITE
    row = (eng.sheet_w + 7) // 8
    assert eng.sheet_bytes == row * eng.sheet_h
    mono, res = eng.process_numpy(pages)
    assert mono.shape == (5, eng.sheet_h, row)
    files, _ = checker.process_sheets_files_cpu(ref_lib, cfg, pages, w, h, U.FMT_GRAY8, out_fmt=U.FMT_MONOBLACK, threads=8)
    for i in range(5):
        rfmt, rw, rh, want = files[i][0]
        assert (rfmt, rw, rh) == (U.FMT_MONOWHITE, eng.sheet_w, eng.sheet_h)
        assert np.array_equal(mono[i], want), f"sheet {i}"
        assert res[i].status == 0
    # device-resident path gives the same bytes
    import torch
    d_in = torch.from_numpy(pages.reshape(-1)).cuda()
    d_out = torch.empty(5 * eng.sheet_bytes, dtype=torch.uint8, device="cuda")
    eng.process_ptr(d_in.data_ptr(), d_out.data_ptr(), 5, False, None)
    assert np.array_equal(d_out.cpu().numpy().reshape(mono.shape), mono)
    # back to the page format
    eng.set_output_format(-1)
    again, _ = eng.process_numpy(pages)
    assert np.array_equal(again, gray)
    eng.close()
    # direct PNM writer (file.c:134-176)
    path = os.path.join(str(tmp_path), "sheet.pbm")
    m0 = np.ascontiguousarray(mono[0])
    rc = L.load().unpaper_b200_write_pnm(path.encode(), m0.ctypes.data, row, eng.sheet_w, eng.sheet_h, U.FMT_MONOWHITE)
    assert rc == 0
    blob = open(path, "rb").read()
    hdr = b"P4\n%d %d\n" % (eng.sheet_w, eng.sheet_h)
    assert blob[:len(hdr)] == hdr and blob[len(hdr):] == m0.tobytes()


def _mono_pages(fmt, w, h, n, seed=200, box=(0.60, 0.72)):
    out = []
    for i in range(n):
        g = synth.gray_page(seed + i, w, h, box=box)
        bits = np.packbits(g < 128, axis=1)
        out.append(bits if fmt == U.FMT_MONOWHITE else bits ^ 0xFF)
    return np.stack(out)


@pytest.mark.parametrize("fmt", [U.FMT_MONOWHITE, U.FMT_MONOBLACK])
def test_engine_mono_pages(ref_lib, fmt):
    """1-bit pages (pbm scans / the PDF path): expanded by the decode stage, processed
    in the 1 B/px working sheet, written as MONOWHITE — against the reference's
    process_sheet() on the same 1-bit pages."""
    from oracle import checker
    from unpaper_gpu_b200.lib import Engine
    w, h = 624, 880
    pages = _mono_pages(fmt, w, h, 5)
    cfg = U.default_sheet_config()
    eng = Engine(cfg, w, h, fmt, group_pages=2, lanes=2)
    assert eng.out_fmt == U.FMT_MONOWHITE and eng.sheet_bytes == (w // 8) * h
    out, res = eng.process_numpy(pages)
    eng.close()
    rout, rres = checker.process_sheets_cpu(ref_lib, "ref_", cfg, pages, w, h, fmt, threads=8)
    if fmt == U.FMT_MONOBLACK:
        rout = rout ^ 0xFF      # the harness hands back MONOBLACK; saveImage() writes MONOWHITE (file.c:205-207)
    for i, (a, b) in enumerate(zip(res, rres)):
        assert a.status == 0 and b.status == 0
        assert a.deskew_mask_count == b.deskew_mask_count
        for k in range(a.deskew_mask_count):
            assert U.rect_tuple(a.deskew_masks[k]) == U.rect_tuple(b.deskew_masks[k])
            assert a.rotation[k] == b.rotation[k]
        for k in range(a.border_count):
            assert U.border_tuple(a.borders[k]) == U.border_tuple(b.borders[k])
    assert np.array_equal(out, rout), f"{int((out != rout).sum())} differing bytes"


def test_engine_mono_a4_properties():
    """Full-size 1-bit A4: deterministic, dark scan edges removed, ink kept."""
    from unpaper_gpu_b200.lib import Engine
    w, h = synth.A4_W, synth.A4_H
    pages = _mono_pages(U.FMT_MONOWHITE, w, h, 3, seed=0, box=(0.76, 0.80))
    eng = Engine(U.default_sheet_config(), w, h, U.FMT_MONOWHITE, group_pages=2, lanes=2)
    out1, res = eng.process_numpy(pages)
    out2, _ = eng.process_numpy(pages)
    eng.close()
    assert np.array_equal(out1, out2)
    for r in res:
        assert r.status == 0 and r.deskew_mask_count == 1
    bits = np.unpackbits(out1, axis=2)
    assert (bits[:, :, :24] == 0).all()               # left scan edge gone
    assert 0.005 < bits.mean() < 0.2


def test_engine_sheet_callback_streams_pnm(tmp_path):
    """The per-sheet completion hook (reference post_process_fn, batch_worker.c:153-158):
    called in sheet order with the finished bytes while later groups are in flight;
    here it writes every sheet with the direct PNM writer."""
    from unpaper_gpu_b200.lib import Engine
    from unpaper_gpu_b200 import lib as L
    w, h = 620, 877
    pages = np.stack([synth.gray_page(i, w, h, box=(0.60, 0.72)) for i in range(7)])
    eng = Engine(U.default_sheet_config(), w, h, U.FMT_GRAY8, group_pages=2, lanes=2)
    plain, _ = eng.process_numpy(pages)
    eng.set_output_format(U.FMT_MONOWHITE)
    row = (w + 7) // 8
    seen = []

    def done(idx, ptr, res):
        seen.append((idx, res.status, res.deskew_mask_count))
        path = os.path.join(str(tmp_path), f"sheet{idx:03d}.pbm").encode()
        return L.load().unpaper_b200_write_pnm(path, ptr, row, w, h, U.FMT_MONOWHITE)

    eng.set_sheet_callback(done)
    mono, res = eng.process_numpy(pages)
    assert [s[0] for s in seen] == list(range(7))
    assert all(s[1] == 0 and s[2] == 1 for s in seen)
    hdr = b"P4\n%d %d\n" % (w, h)
    for i in range(7):
        blob = open(os.path.join(str(tmp_path), f"sheet{i:03d}.pbm"), "rb").read()
        assert blob == hdr + mono[i].tobytes()
    # a failing hook fails the call (after draining), and the engine stays usable
    eng.set_sheet_callback(lambda idx, ptr, res: 1 if idx == 3 else 0)
    with pytest.raises(RuntimeError, match="-3"):
        eng.process_numpy(pages)
    eng.set_sheet_callback(None)
    eng.set_output_format(-1)
    again, _ = eng.process_numpy(pages)
    assert np.array_equal(again, plain)
    eng.close()


@pytest.mark.parametrize("page_fmt,out_fmt", [(U.FMT_GRAY8, -1), (U.FMT_GRAY8, U.FMT_MONOWHITE), (U.FMT_RGB24, -1),
                                              (U.FMT_RGB24, U.FMT_GRAY8), (U.FMT_MONOBLACK, -1)])
def test_engine_output_split(ref_lib, page_fmt, out_fmt):
    """--output-pages 2 (sheet_stages.c:606-621): the double-layout sheet leaves as two
    images of width sheet_w/2, each converted like saveImage() does — against the files the
    reference's own output stage writes.  Odd sheet width: the last column is dropped."""
    from oracle import checker
    from unpaper_gpu_b200.lib import Engine
    w, h = 1241, 620
    if page_fmt == U.FMT_RGB24:
        pages = np.stack([np.concatenate([synth.color_page(300 + 2 * i + k, 620 + k, h) .reshape(h, 620 + k, 3) for k in (0, 1)], axis=1).reshape(h, -1)
                          for i in range(3)])
    else:
        g = np.stack([synth.double_sheet(300 + i, w, h) for i in range(3)])
        pages = g if page_fmt == U.FMT_GRAY8 else np.stack([np.packbits(x >= 128, axis=1) for x in g])
    cfg = U.default_sheet_config()
    cfg.layout, cfg.output_count = U.LAYOUT_DOUBLE, 2
    if page_fmt == U.FMT_RGB24:
        cfg.no_blackfilter = cfg.no_noisefilter = 1
    eng = Engine(cfg, w, h, page_fmt, group_pages=2, lanes=2)
    if out_fmt >= 0:
        eng.set_output_format(out_fmt)
    assert eng.out_count == 2 and eng.out_w == w // 2
    out, res = eng.process_numpy(pages)
    import torch
    d_in = torch.from_numpy(np.ascontiguousarray(pages).reshape(-1)).cuda()
    d_out = torch.empty(out.size, dtype=torch.uint8, device="cuda")
    eng.process_ptr(d_in.data_ptr(), d_out.data_ptr(), 3, False, None)
    eng.close()
    assert np.array_equal(d_out.cpu().numpy().reshape(out.shape), out)
    files, rres = checker.process_sheets_files_cpu(ref_lib, cfg, pages, w, h, page_fmt, out_fmt=out_fmt, output_count=2, threads=8)
    for i in range(3):
        assert res[i].status == 0 and rres[i].status == 0
        for j in range(2):
            rfmt, rw, rh, want = files[i][j]
            assert (rfmt, rw, rh) == (eng.out_fmt, w // 2, h)
            assert np.array_equal(out[i, j], want), f"sheet {i} page {j}: {int((out[i, j] != want).sum())} differing bytes"
