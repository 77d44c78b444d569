"""states written by gravity_scenario_save must be read
back by the REFERENCE's own parser (sim_gravity.c:272-402, run here through
oracle/_ref/ref_gravity, the unmodified translation unit) as the same bits."""
import os

import numpy as np
import pytest

import proj_entropy_b200 as g
from conftest import ROOT, bits_equal, hexarr, load_fixture, ok_fixture_names, scenario_from_fixture
from oracle import oracle_py as oracle

needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/ref_gravity not built")


def reference_parse(path):
    rc, lines = oracle.ref_run(path, [0])
    assert lines, "no output from the reference harness"
    return rc, lines[0]


def assert_same_state(parsed, names, dt, sim_width, mass, radius, x, y, vx, vy):
    assert parsed["n"] == len(names) and parsed["dt"] == dt
    assert float.fromhex(parsed["sim_width"]) == sim_width
    assert [b["name"] for b in parsed["bodies"]] == list(names)
    for key, want in (("mass", mass), ("radius", radius), ("x", x), ("y", y), ("vx", vx), ("vy", vy)):
        got = hexarr([b[key] for b in parsed["bodies"]])
        assert bits_equal(got, want), key


@needs_ref
def test_plummer_sample_with_empty_names_is_read_by_the_reference(tmp_path):
    """The case the row names: a <= 200-body Plummer sample has no names at all."""
    n = 200
    m, x, y, vx, vy = g.plummer(n, seed=7)
    sc = g.Scenario("PLUMMER_200", 1, 50 * 1.495978707E11, [""] * n, m, np.full(n, 1.0E6), x, y, vx, vy,
                    np.full(n, 9, np.int32), np.full(n, 30, np.int32))
    path = str(tmp_path / "plummer_200")
    sc.save(path)
    rc, parsed = reference_parse(path)
    assert rc == 0, parsed
    assert_same_state(parsed, ["B%d" % i for i in range(n)], 1, 50 * 1.495978707E11, m, np.full(n, 1.0E6), x, y, vx, vy)
    # and our own reader agrees with the reference's
    again = g.load_scenario(path)
    assert again.names[:2] == ["B0", "B1"] and bits_equal(again.x, x) and bits_equal(again.vy, vy)


@needs_ref
@pytest.mark.parametrize("name", ok_fixture_names())
def test_every_fixture_state_round_trips_through_the_reference_parser(name, tmp_path):
    """Nested scenarios (relative coordinates resolved), DT = 0 (no PARAM line: inherit),
    every SIM_WIDTH of the fixtures, 15 bodies ... saved flat, parsed by the reference."""
    fx = load_fixture(name)
    sc = scenario_from_fixture(fx)
    path = str(tmp_path / name)
    sc.save(path)
    text = open(path).read()
    assert ("DT_LINUX" in text) == (fx["dt"] != 0)
    rc, parsed = reference_parse(path)
    assert rc == 0, parsed
    cp = fx["checkpoints"][0]
    assert_same_state(parsed, fx["names"], fx["dt"], float.fromhex(fx["sim_width"]), fx["mass_f"], hexarr(fx["radius"]),
                      cp["x_f"], cp["y_f"], cp["vx_f"], cp["vy_f"])


@needs_ref
def test_every_sim_width_of_the_table_survives_the_snap(tmp_path):
    """SIM_WIDTH is snapped DOWN on load (sim_gravity.c:299-311): value/AU*AU one ulp short
    would fall to the next entry.  All 27 table entries, and values between entries."""
    au = 1.495978707E11
    lits = [.01, .02, .05, .1, .2, .5, 1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000, 50000,
            100000, 200000, 500000, 1000000, 2000000, 5000000]
    m, x, y, vx, vy = g.plummer(3, seed=1)
    for k, lit in enumerate(lits):
        for width, want in ((lit * au, lit * au), (lit * au * 1.3, lit * au)):
            sc = g.Scenario("W", 1, width, ["a", "b", "c"], m, np.ones(3), x, y, vx, vy, np.zeros(3, np.int32),
                            np.zeros(3, np.int32))
            path = str(tmp_path / ("w%d" % k))
            sc.save(path)
            rc, parsed = reference_parse(path)
            assert rc == 0 and float.fromhex(parsed["sim_width"]) == want, (lit, width)


@needs_ref
def test_awkward_names_become_single_tokens(tmp_path):
    names = ["", "   ", "two words", "#hash", "PARAM", "tab\there", "ok", "x" * 31]
    n = len(names)
    m, x, y, vx, vy = g.plummer(n, seed=3)
    sc = g.Scenario("NAMES", 60, 3 * 1.495978707E11, names, m, np.ones(n), x, y, vx, vy, np.zeros(n, np.int32),
                    np.zeros(n, np.int32))
    path = str(tmp_path / "names")
    sc.save(path)
    rc, parsed = reference_parse(path)
    assert rc == 0, parsed
    got = [b["name"] for b in parsed["bodies"]]
    assert got == ["B0", "B1", "two_words", "_#hash", "_PARAM", "tab_here", "ok", "x" * 31]
    assert bits_equal(hexarr([b["x"] for b in parsed["bodies"]]), x)


def test_states_the_format_cannot_carry_are_refused(tmp_path):
    m, x, y, vx, vy = g.plummer(201, seed=1)
    mk = lambda n, dt, xs=None: g.Scenario("S", dt, 3 * 1.495978707E11, ["b%d" % i for i in range(n)], m[:n], np.ones(n),
                                           x[:n] if xs is None else xs, y[:n], vx[:n], vy[:n],
                                           np.zeros(n, np.int32), np.zeros(n, np.int32))
    with pytest.raises(ValueError, match="200"):
        mk(201, 1).save(str(tmp_path / "big"))              # the reference stops reading at MAX_OBJECT
    mk(200, 1).save(str(tmp_path / "ok"))
    with pytest.raises(ValueError, match="DT"):
        mk(5, 7000).save(str(tmp_path / "dt"))              # would be snapped to 3600*... on load
    bad = x[:5].copy()
    bad[2] = np.nan
    with pytest.raises(ValueError, match="finite"):
        mk(5, 1, bad).save(str(tmp_path / "nan"))
    with pytest.raises(OSError):
        mk(5, 1).save(str(tmp_path / "no_such_dir" / "f"))
    assert not os.path.exists(str(tmp_path / "big"))
