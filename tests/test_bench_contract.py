"""bench.py's output contract (task prompt, "Measurement"): the committed bench line of the round
(profiles/r02_bench_n1.json, produced on a B200 by tools/profile_round.sh) and a live `--impl reference` line carry
every key the driver reads, with consistent values."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"]


def check_common(line):
    for k in BASE_KEYS:
        assert k in line, k
    assert line["metric"].startswith("Mtri/s") and line["unit"] == "Mtri/s" and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["vs_baseline"] is None and line["data"] == "synthetic"   # a fixed batch of 64 views dealt to the ranks
    assert "workload" in line["config"] and "model" not in line["config"]
    e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["unit"] == line["unit"]
    c = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1


COMMITTED = os.path.join(ROOT, "profiles", "r02_bench_n1.json")


@pytest.mark.skipif(not os.path.exists(COMMITTED), reason="no committed bench line of this round yet")
def test_committed_gpu_bench_line():
    line = json.load(open(COMMITTED))
    check_common(line)
    cfg = line["config"]
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["gpu_launches"] >= line["steps"] * 64 * 4
    assert cfg["triangles_per_step"] == 64 * cfg["triangles_per_view"]
    assert abs(line["value"] - cfg["triangles_per_step"] / (line["ms_per_step"] * 1e-3) / 1e6) < 0.01 * line["value"]
    assert line["rasterized_Mtri_s"] < line["processed_Mtri_s"] < line["value"]
    clocks = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clocks)
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] > r["algorithmic_bytes"] * 0.5
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        assert abs(r["peak"] - json.load(open(peaks))["hbm_gbs"]) < 1e-6, "roofline.peak must be the driver-measured HBM rate"
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] >= 1728 * cfg["meshlets"] and e["d2h_bytes_per_step"] == 64 * 1920 * 1080 * 4
    assert e["value"] < line["value"]                      # copies inside the timed region can only cost
    assert line["cpu_baseline"]["value"] < e["value"]
    assert "l2" in cfg and "timing" in cfg
    p = line["parity"]
    assert p["views_checked"] == 64 and p["visbuffer_exact"] == 64 and p["colour_ok"] == p["colour_checked"] >= 1
    assert line["draw_stats"]["records"] > 0               # the binner and the tile rasterizer have work in the headline
    for name in ("c1_knot", "c1_sponza", "c2_grid", "c3_knot", "c5_views"):
        for v in line["configs"][name]["views"]:
            assert v["parity"]["visbuffer"] == "exact", name


def test_reference_arm_line_live():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()
    assert len(out) == 1, "exactly one line on stdout"
    line = json.loads(out[0])
    check_common(line)
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]
    if os.path.exists(COMMITTED):
        ours = json.load(open(COMMITTED))
        assert line["metric"] == ours["metric"] and line["config"]["workload"] == ours["config"]["workload"]
        for k in ("step", "triangles_per_view", "triangles_per_step", "meshlets", "draws_per_view"):
            assert line["config"][k] == ours["config"][k], k
