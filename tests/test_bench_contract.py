"""The JSON contract of bench.py, checked offline on the lines recorded in
profiles/ (produced on a B200 by the commands named in profiles/README.md), and
the CPU-only --impl reference arm run live."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def check_line(d, reference=False):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "fp64 pair-interactions/s" and d["unit"] == "pairs/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    cb = d.get("cpu_baseline")
    if reference:
        assert d["impl"] == "reference" and d["gpu_launches"] == 0
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] == cb["value"]
    if cb:
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]


def test_recorded_gpu_line_meets_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_1gpu_final.json")))
    check_line(d)
    n = d["config"]["bodies"]
    assert abs(d["value"] - n * (n - 1) / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
    assert d["gpu_launches"] == d["steps"]                       # one kernel per step
    assert d["e2e"]["h2d_bytes_per_step"] == 5 * 8 * n and d["e2e"]["d2h_bytes_per_step"] == 4 * 8 * n
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.01
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.5 < r["frac"] < 0.8
    assert abs(r["achieved"] - 20 * d["value"] / 1e12) / r["achieved"] < 1e-3
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                       "sw_thermal_slowdown"}
    assert d["cpu_baseline"]["value"] < d["value"] / 50


def test_recorded_multi_gpu_lines():
    for name, gpus in (("r1_scale_g4_n1048576.json", 4), ("r1_scale_g8_n1048576.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        check_line(d)
        assert d["n_gpus"] == gpus and d["scaling"] == "strong" and "cpu_baseline" not in d
        assert d["e2e"]["h2d_bytes_per_step"] == 5 * 8 * d["config"]["bodies"] * gpus


def test_reference_arm_live():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--bodies", "8192", "--cpu-rows", "1024"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    check_line(json.loads(lines[0]), reference=True)
