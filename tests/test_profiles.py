"""The committed profile evidence is what the tools make of the committed ncu export (CPU only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_json_is_the_capture(tmp_path):
    """profiles/traffic.json (the source of `roofline.traffic` in the bench line) regenerated from the raw
    page of the ncu capture it names."""
    committed = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    src = next(iter(committed.values()))["source"]
    out = tmp_path / "traffic.json"
    with open(os.path.join(ROOT, "profiles", "r02_prof_r02d_raw.csv")) as f:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stage_traffic.py"), "prof_r02d", "32", src, str(out)],
                       stdin=f, stdout=subprocess.DEVNULL, check=True)
    assert json.load(open(out)) == committed


def test_traffic_stages_are_bench_stages():
    sys.path.insert(0, ROOT)
    import bench
    stages = set(bench.stage_bytes(1.0, 1.0))
    committed = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert set(committed) <= stages
    # every stage that moves sheet bytes has a DRAM figure; deskew (the dominant stage of the bench line) above all
    assert {"decode", "blackfilter", "noisefilter", "blurfilter", "grayfilter", "detect_masks", "detect_rotation",
            "deskew", "center_mask", "border"} <= set(committed)
    S = 2480 * 3508
    assert 1.0 * S < committed["deskew"]["bytes_per_sheet"] < 2.2 * S      # read + write of the sheet, no re-reads
