"""Regenerates tests/golden/golden_vectors.json from the UNMODIFIED reference
(oracle/_ref built from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

Vectors:
  * "A1": the reference's own end-to-end golden (tests/unpaper_tests.py:653-669):
    default pipeline on tests/source_images/imgsrc001.png; stores what the
    reference CPU backend decided, a digest of its output, and the differing-
    pixel ratio against tests/golden_images/goldenA1.pbm (pins oracle/_ref to
    the reference's golden image);
  * "sheets": seeded synthetic pages (generator in unpaper-gpu_b200/synth.py)
    through the reference process_sheet(): decisions + output digests;
  * "ops": digests of single reference ops on seeded images.
Nothing here reads /root/reference at test time; only this script does.
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import unpaper_gpu_b200 as U  # noqa: E402
from unpaper_gpu_b200 import synth  # noqa: E402
from oracle import checker  # test infrastructure: the CPU checkers
import golden_cases as G  # noqa: E402


def main():
    lib = checker.load_ref()
    assert lib is not None, "build oracle/_ref first (make -C oracle ref)"
    ops = U.HostOps(lib, "ref_host_")
    out = {"A1": None, "sheets": {}, "ops": {}}
    ref_root = os.environ.get("UNPAPER_REFERENCE", "/root/reference")
    src = os.path.join(ref_root, "tests/source_images/imgsrc001.png")
    if os.path.exists(src):
        from PIL import Image
        g = np.array(Image.open(src).convert("L"), dtype=np.uint8)
        cfg = U.default_sheet_config()
        o, res = checker.process_sheets_cpu(lib, "ref_", cfg, g, g.shape[1], g.shape[0], U.FMT_GRAY8)
        gold = np.array(Image.open(os.path.join(ref_root, "tests/golden_images/goldenA1.pbm")).convert("L"))
        ratio = float(np.mean((o[0] < cfg.abs_black_threshold) != (gold < 128)))
        out["A1"] = {"input_sha256": hashlib.sha256(g.tobytes()).hexdigest(), "size": [int(g.shape[1]), int(g.shape[0])],
                     "result": G.result_dict(res[0]), "output_sha256": hashlib.sha256(o[0].tobytes()).hexdigest(),
                     "golden_diff_ratio_thr170": ratio}
        assert ratio < 1e-4, ratio
    for name, (cfg, pages, w, h, fmt) in G.sheet_cases().items():
        o, res = checker.process_sheets_cpu(lib, "ref_", cfg, pages, w, h, fmt, threads=8)
        out["sheets"][name] = {"results": [G.result_dict(r) for r in res],
                               "output_sha256": [hashlib.sha256(x.tobytes()).hexdigest() for x in o]}
    for name, fn in G.op_cases().items():
        out["ops"][name] = fn(ops)
    with open(os.path.join(HERE, "golden_vectors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote golden_vectors.json:", {k: (len(v) if v else 0) for k, v in out.items()})


if __name__ == "__main__":
    main()
