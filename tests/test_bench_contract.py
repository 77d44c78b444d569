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
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_1gpu.json")))
    check_line(d)
    n = d["config"]["bodies"]
    assert abs(d["value"] - n * (n - 1) / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
    assert d["gpu_launches"] == 1 and d["variant"]["steps_per_launch"] == d["steps"]     # the whole batch in one launch
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
    check_parity(d["parity"], 1)
    # both arms name the workload identically (the driver compares them)
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_reference_arm.json")))
    check_line(ref, reference=True)
    assert ref["config"] == d["config"]
    assert ref["full_step_ms_extrapolated"] > 100 * ref["ms_per_step"]       # ms_per_step is what was really timed


def check_parity(p, world):
    """The in-run parity record every line carries."""
    assert p["rows"] >= 512 and p["ranks_compared"] == world and p["ranks_bit_identical"] is True
    assert p["max_rel_vs_ref_with_floor"] <= 1e-12 and p["n_above_1e-12"] <= 1 and p["pass"] is True
    assert p["floor_used"] == (p["n_above_1e-12"] > 0) and "1e-12" in p["gate"]
    assert p["max_rel_vs_long_double"] <= p["ref_max_rel_vs_long_double"]   # no farther from truth than the reference


def test_recorded_multi_gpu_lines():
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_scale_g*_n*_x*.json")))
    assert len(paths) == 21                                     # 3 sizes x (1 GPU + {2,4,8} GPUs x 2 exchanges)
    for path in paths:
        d = json.load(open(path))
        check_line(d)
        gpus, n = d["n_gpus"], d["config"]["bodies"]
        assert "_g%d_n%d_x%d" % (gpus, n, d["variant"]["exchange"] if gpus > 1 else 1) in path
        assert d["scaling"] == "strong" and "cpu_baseline" not in d
        # every rank copies only its own shard: totals do not grow with the number of GPUs
        assert d["e2e"]["h2d_bytes_per_step"] == 5 * 8 * n and d["e2e"]["d2h_bytes_per_step"] == 4 * 8 * n
        check_parity(d["parity"], gpus)
        # north_star's bar (>= 60 % of aggregate nominal FP64 peak), every cell with the fused
        # peer-memory exchange; the NCCL all-gather variant may sit a hair below it at 8 x 64K
        bar = 0.60 if (gpus == 1 or d["variant"]["exchange"] == 1) else 0.595
        assert d["roofline"]["job_frac_of_nominal"] >= bar, path


def test_reference_arm_live():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--bodies", "8192", "--cpu-rows", "1024"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    check_line(json.loads(lines[0]), reference=True)
