"""The measurement evidence under profiles/ is self-consistent (no GPU needed): what bench.py reports as roofline.traffic is
what the committed raw ncu export of the shipped force kernel says, and the launch summary is derived from the launch list."""
import importlib.util
import json
import os

from conftest import ROOT


def _tool():
    spec = importlib.util.spec_from_file_location("ncu_summary", os.path.join(ROOT, "tools", "ncu_summary.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_roofline_traffic_matches_the_raw_ncu_export():
    tab = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    rec = tab["256_ml32_n1"]
    got = _tool().traffic(os.path.join(ROOT, "profiles", "r2_ncu_rows2_256_raw.csv"), "256_ml32_n1", "p2p_rows2", write=False)
    assert got["dram_bytes"] == rec["dram_bytes"] == got["dram_read"] + got["dram_write"]
    assert got["registers"] == rec["registers"] and got["duration_ms"] == rec["duration_ms"]
    # the capture is of the shipped build: one pass per row, no blocked summation, no retiring warps
    raw = open(os.path.join(ROOT, "profiles", "r2_ncu_rows2_256_raw.csv")).read()
    assert "p2p_rows2_kernel<1, 384, 1, 3, 0, 0, 0>" in raw
    # algorithmic bytes of the launch (32 B per particle + 4 B per task at 256^3) against the measured traffic: no wasted re-reads
    algorithmic = 16777216 * 32 + 105453924 * 4
    assert algorithmic < rec["dram_bytes"] < 1.6 * algorithmic


def test_launch_summary_is_derived_from_the_launch_list(tmp_path):
    out = tmp_path / "summary.txt"
    _tool().launches(os.path.join(ROOT, "profiles", "r2_launches_bench256_n1.csv"), str(out))
    assert out.read_text() == open(os.path.join(ROOT, "profiles", "r2_launches_bench256_n1_summary.txt")).read()
    first = out.read_text().splitlines()[1]
    assert "p2p_rows2_kernel" in first and float(first.split()[-2]) > 60.0        # the force kernel dominates the step
