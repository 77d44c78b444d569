"""The product's scenario reader (libgravity_host, gravity_scenario_load) against
what the reference's own parser (sim_gravity_init, sim_gravity.c:197-422)
produced for the same file, recorded in tests/golden/ by the unmodified TU."""
import hashlib
import json
import os

import numpy as np
import pytest

import proj_entropy_b200 as g
from conftest import GOLDEN, bits_equal, fixture_names, load_fixture, scenario_path


def _cases():
    out = []
    for name in fixture_names():
        fx = json.load(open(os.path.join(GOLDEN, name + ".json")))
        out.append(pytest.param(name, id=name))
    return out


@pytest.mark.parametrize("name", _cases())
def test_reader_matches_reference_parser(name):
    fx = load_fixture(name)
    path = scenario_path(fx)
    if path is None:
        pytest.skip("scenario text lives in /root/reference (build container only)")
    assert hashlib.md5(open(path, "rb").read()).hexdigest() == fx["file_md5"]
    if "error" in fx:
        with pytest.raises(g.ScenarioError) as ei:
            g.load_scenario(path)
        rows = ei.value.rows
        assert rows[0] == fx["error"][0]
        assert os.path.basename(rows[1]) == fx["error"][1]
        assert rows[2] == fx["error"][2]
        return
    s = g.load_scenario(path)
    cp0 = fx["checkpoints"][0]
    assert s.n == fx["n"]
    assert s.dt == fx["dt"]
    assert s.sim_width == float.fromhex(fx["sim_width"])
    assert s.names == fx["names"]
    assert s.sim_name == os.path.basename(path).upper()[:31]
    assert bits_equal(s.mass, fx["mass_f"])
    assert bits_equal(s.radius, [float.fromhex(v) for v in fx["radius"]])
    for got, key in ((s.x, "x_f"), (s.y, "y_f"), (s.vx, "vx_f"), (s.vy, "vy_f")):
        assert bits_equal(got, cp0[key]), key


def test_open_failed_rows():
    with pytest.raises(g.ScenarioError) as ei:
        g.load_scenario("/nonexistent/scenario")
    assert ei.value.rows[0] == "Open Failed" and ei.value.rows[1] == "/nonexistent/scenario"


def test_empty_file_fails_to_open(tmp_path):
    p = tmp_path / "empty"
    p.write_text("")
    with pytest.raises(g.ScenarioError) as ei:
        g.load_scenario(str(p))
    assert ei.value.rows[0] == "Open Failed"      # util.c:1761-1765: a zero-length read is an error


def test_params_are_inherited_when_absent():
    fx = load_fixture("own_no_params")
    path = scenario_path(fx)
    assert g.load_scenario(path).dt == 0 == fx["dt"]          # fresh process: statics are zero
    s = g.load_scenario(path, inherit_dt=30, inherit_sim_width=5.0)
    assert s.dt == 30 and s.sim_width == 5.0                  # sim_gravity.c:222-223


def test_object_cap(tmp_path):
    p = tmp_path / "many"
    p.write_text("PARAM DT_LINUX 1 SEC\n" + "".join("b%d %d 0 0 0 1E20 1 RED 1\n" % (i, i * 1000) for i in range(250)))
    assert g.load_scenario(str(p)).n == 200                   # MAX_OBJECT, sim_gravity.c:398-401
    assert g.load_scenario(str(p), max_objects=1000).n == 250  # extension for large hosts


def test_tab_only_line_is_not_blank(tmp_path):
    # only ' ' counts as blank (sim_gravity.c:280-285); a tab goes on to the
    # object scan and fails it
    p = tmp_path / "tab"
    p.write_text("a 0 0 0 0 1E30 1 RED 1\n\t\n")
    with pytest.raises(g.ScenarioError) as ei:
        g.load_scenario(str(p))
    assert ei.value.rows[0] == "Invalid Object" and ei.value.rows[2] == "line 2"


def test_save_round_trip(tmp_path):
    fx = load_fixture("own_nested_indent")
    s = g.load_scenario(scenario_path(fx))
    out = tmp_path / "saved"
    s.save(str(out))
    t = g.load_scenario(str(out))
    assert t.n == s.n and t.dt == s.dt and t.sim_width == s.sim_width and t.names == s.names
    for k in ("mass", "radius", "x", "y", "vx", "vy"):
        assert bits_equal(getattr(s, k), getattr(t, k)), k      # %.17g round-trips doubles


def test_snapshot_round_trip(tmp_path):
    m, x, y, vx, vy = g.plummer(1234, seed=3)
    p = tmp_path / "snap.bin"
    g.snapshot_write(str(p), 86399, 7, 0, m, x, y, vx, vy)
    s = g.snapshot_read(str(p))
    assert (s["n"], s["sim_time"], s["dt"], s["last_day"]) == (1234, 86399, 7, 0)
    for a, b in zip((m, x, y, vx, vy), (s["mass"], s["x"], s["y"], s["vx"], s["vy"])):
        assert bits_equal(a, b)
    assert os.path.getsize(p) == 8 + 8 + 8 + 4 + 4 + 8 + 8 + 5 * 8 * 1234 and s["path_tail"] == 0 and len(s["path_xy"]) == 0
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"NOTASNAP" + bytes(64))
    with pytest.raises(OSError):
        g.snapshot_read(str(bad))
    # version 1 files (no PATH section) are still read
    import struct
    v1 = tmp_path / "v1.bin"
    v1.write_bytes(b"GRAVSNP1" + struct.pack("<qqii", 3, 500, 60, 2) + np.arange(15, dtype=np.float64).tobytes())
    s = g.snapshot_read(str(v1))
    assert (s["n"], s["sim_time"], s["dt"], s["last_day"], s["path_tail"]) == (3, 500, 60, 2, 0)
    assert s["mass"].tolist() == [0.0, 1.0, 2.0] and s["vy"].tolist() == [12.0, 13.0, 14.0]


def test_snapshot_carries_the_path_rings(tmp_path):
    """Version 2: the samples taken so far and the most recent ones, oldest first, so that a resumed run
    continues the PATH rings (sim_gravity.c:1248-1260); truncated and inconsistent files are refused."""
    n = 7
    m, x, y, vx, vy = g.plummer(n, seed=5)
    rings = np.random.default_rng(1).normal(size=(40, n, 2))
    p = tmp_path / "snap2.bin"
    g.snapshot_write(str(p), 86400 * 123, 86400, 122, m, x, y, vx, vy, path_tail=100040, path_xy=rings)
    s = g.snapshot_read(str(p))
    assert (s["n"], s["sim_time"], s["dt"], s["last_day"], s["path_tail"]) == (n, 86400 * 123, 86400, 122, 100040)
    assert bits_equal(s["path_xy"], rings) and bits_equal(s["x"], x)
    raw = p.read_bytes()
    (tmp_path / "cut.bin").write_bytes(raw[:-8])
    with pytest.raises(OSError):
        g.snapshot_read(str(tmp_path / "cut.bin"))
    with pytest.raises(OSError):                       # more samples stored than ever taken
        g.snapshot_write(str(tmp_path / "no.bin"), 0, 1, 0, m, x, y, vx, vy, path_tail=3, path_xy=rings)


def test_out_of_range_dt_behaves_like_the_reference_build(tmp_path):
    # (int32_t)1e300 and (int32_t)NaN are undefined in C; the reference's x86-64
    # build produces INT32_MIN, which snaps to the table floor of 1 second
    for text in ("1e300", "nan", "-1e12", "4e9"):
        p = tmp_path / "dt"
        p.write_text("PARAM DT_LINUX %s SEC\na 0 0 0 0 1E30 1 RED 5\n" % text)
        assert g.load_scenario(str(p)).dt == 1, text
