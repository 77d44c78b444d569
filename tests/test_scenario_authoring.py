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


@needs_ref
def test_writer_fuzz_against_the_reference_parser(tmp_path):
    """80 seeded random states -- 1 to 200 bodies, values from subnormal to 1e300, negative
    zeros excepted (the format turns -0.0 into +0.0), random 7-bit bytes (blanks, tabs, control
    characters, '#') in the names, every DT of
    the table and 0, random widths -- written by gravity_scenario_save and parsed back by the
    unmodified reference: same count, DT, snapped width, and every double bit for bit."""
    import random
    rng = random.Random(20261020)
    valid_dt = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 20, 30, 40, 50, 60, 120, 180, 240, 300, 360, 420, 480, 540, 600, 1200, 1800,
                2400, 3000, 3600, 7200, 10800, 14400, 18000, 21600, 86400]
    au = 1.495978707E11
    table = [v * au for v in (.01, .02, .05, .1, .2, .5, 1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000,
                              20000, 50000, 100000, 200000, 500000, 1000000, 2000000, 5000000)]

    def value():
        k = rng.random()
        if k < 0.1:
            return 0.0
        if k < 0.15:
            return rng.choice([5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1e300, -1e300])
        return rng.uniform(-1, 1) * 10.0 ** rng.randint(-30, 30)

    def name():
        k = rng.random()
        if k < 0.15:
            return ""
        # any 7-bit byte except the two the harness's JSON output cannot carry unescaped
        raw = bytes(rng.choice([c for c in range(1, 128) if c not in (0x22, 0x5c)]) for _ in range(rng.randint(1, 31)))
        return raw.decode("latin-1")

    for case in range(80):
        n = rng.choice([1, 2, 3, 7, 15, 16, 50, 199, 200])
        cols = [np.array([value() for _ in range(n)]) for _ in range(6)]      # mass radius x y vx vy
        names = [name() for _ in range(n)]
        dt = rng.choice(valid_dt)
        width = rng.choice(table) * rng.choice([1.0, 1.0, 1.7, 0.999999999])
        sc = g.Scenario("FUZZ", dt, width, names, cols[0], cols[1], cols[2], cols[3], cols[4], cols[5],
                        np.array([rng.randrange(10) for _ in range(n)], np.int32),
                        np.array([rng.randrange(0, 5000) for _ in range(n)], np.int32))
        # the mirror passes names as latin-1 bytes
        sc.names = [nm.encode("latin-1").decode("latin-1") for nm in names]
        path = str(tmp_path / ("fuzz%d" % case))
        _save_latin1(sc, path)
        rc, lines = oracle.ref_run(path, [0])
        assert rc == 0 and lines, (case, lines)
        parsed = lines[0]
        assert parsed["n"] == n and parsed["dt"] == dt, case
        want_width = max([t for t in table if t <= width] or [table[0]])
        assert float.fromhex(parsed["sim_width"]) == want_width, (case, width)
        for key, col in zip(("mass", "radius", "x", "y", "vx", "vy"), cols):
            got = hexarr([b[key] for b in parsed["bodies"]])
            want = col + 0.0                                                       # -0.0 -> +0.0 is the format's, not ours
            assert bits_equal(got, want), (case, key)
        # names came back as single tokens, one per body, none empty
        assert all(len(b["name"]) >= 1 and not any(c in b["name"] for c in " \t\n\r\x0b\x0c") for b in parsed["bodies"])
        # and our own reader agrees with the reference's
        again = g.load_scenario(path)
        assert again.n == n and again.dt == dt and bits_equal(again.x, cols[2] + 0.0) and bits_equal(again.mass, cols[0] + 0.0)


def _save_latin1(sc, path):
    """Scenario.save encodes names as UTF-8; the fuzz wants arbitrary bytes, so build the C struct here."""
    import ctypes as C
    from proj_entropy_b200 import host
    lib = host.load_host_library()
    n = sc.n
    names = (C.c_char * 32 * n)()
    for i, nm in enumerate(sc.names):
        raw = nm.encode("latin-1")[:31]
        C.memmove(names[i], raw, len(raw))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (sc.mass, sc.radius, sc.x, sc.y, sc.vx, sc.vy)]
    col = np.ascontiguousarray(sc.colors, dtype=np.int32)
    pd = np.ascontiguousarray(sc.max_path_disp_dflt, dtype=np.int32)
    cs = host._CScenario()
    cs.sim_name = b"FUZZ"
    cs.sim_time, cs.dt, cs.n, cs.sim_width = 0, int(sc.dt), n, float(sc.sim_width)
    cs.name = C.cast(names, C.POINTER(C.c_char * 32))
    cs.mass, cs.radius, cs.x, cs.y, cs.vx, cs.vy = [a.ctypes.data_as(dp) for a in arrs]
    cs.disp_color, cs.max_path_disp_dflt = col.ctypes.data_as(ip), pd.ctypes.data_as(ip)
    assert lib.gravity_scenario_save(C.byref(cs), os.fsencode(path)) == 0
