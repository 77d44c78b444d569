"""Parity of the CUDA path (through the C ABI of libgravity_cuda) against the
pinned oracle and the golden fixtures.  Everything here needs a B200.

Bars (BASELINE.json north_star):
  * STRICT kernels: bit-identical to the reference;
  * TILED kernel: per-step accelerations <= 1e-12 relative, trajectories
    <= 1e-9 relative after 10^4 steps on the non-chaotic scenarios, energy
    drift within the same tolerance.
"""
import ctypes

import numpy as np
import pytest

import proj_entropy_b200 as g
from conftest import assert_accel_parity, bits_equal, load_fixture, ok_fixture_names, rel_err_rows
from oracle import oracle_py as oracle

pytestmark = pytest.mark.gpu

ACCEL_TOL = 1e-12
TRAJ_TOL = 1e-9
NTHREADS = 16


def state0(fx):
    cp = fx["checkpoints"][0]
    return fx["mass_f"], cp["x_f"], cp["y_f"], cp["vx_f"], cp["vy_f"]


# ---------------------------------------------------------------------------
# STRICT: bit-identical to the unmodified reference
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("name", ok_fixture_names())
def test_strict_small_kernel_reproduces_reference_bits(name):
    """All scenarios, every checkpoint, including 10^5 steps of
    solar_system_and_rogue_star (BASELINE.json configs[1])."""
    fx = load_fixture(name)
    with g.GravityCuda(fx["n"]) as h:            # AUTO -> STRICT for N <= 1024
        h.upload(*state0(fx))
        done = 0
        for cp in fx["checkpoints"][1:]:
            h.step(fx["dt"], cp["steps"] - done)
            done = cp["steps"]
            x, y, vx, vy = h.download()
            for got, key in ((x, "x_f"), (y, "y_f"), (vx, "vx_f"), (vy, "vy_f")):
                assert bits_equal(got, cp[key]), "%s step %d %s" % (name, done, key)


@pytest.mark.parametrize("n", [1, 2, 32, 33, 129, 200, 777, 1024, 1500, 3000])
def test_strict_kernels_bits_vs_oracle(n):
    """Every strict kernel shape: pair-parallel (N <= 32), cooperative multi-step
    grid (N <= 1024), rows (any N); odd and even step batches so that the state
    ends in either ping-pong buffer."""
    m, x, y, vx, vy = g.plummer(n, seed=3)
    with g.GravityCuda(n, kernel=g.KERNEL_STRICT) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(7)
        rax, ray = oracle.accel(m, x, y, vx, vy, 7, nthreads=NTHREADS)
        assert bits_equal(ax, rax) and bits_equal(ay, ray)
        h.step(7, 3)
        ax, ay = h.accel(7)
        h.step(7, 2)
        h.step(7, 1)
        got = h.download()
    mid = oracle.step(m, x, y, vx, vy, 7, 3, nthreads=NTHREADS)
    rax, ray = oracle.accel(m, *mid, 7, nthreads=NTHREADS)
    assert bits_equal(ax, rax) and bits_equal(ay, ray)
    want = oracle.step(m, *mid, 7, 3, nthreads=NTHREADS)
    for a, b in zip(got, want):
        assert bits_equal(a, b)


def test_strict_large_shard_takes_the_row_kernel_and_stays_bit_exact():
    """Above 65,536 bodies per device STRICT mode switches from the term-parallel kernel to one
    thread per body; both must give the reference's bits (accelerations and one step)."""
    n = 66000
    m, x, y, vx, vy = g.plummer(n, seed=23)
    x[500], y[500] = x[64000], y[64000]                    # a coincident pair across the two halves
    vx[64000] = vy[64000] = 0.0
    with g.GravityCuda(n, kernel=g.KERNEL_STRICT) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(3)
        h.step(3, 1)
        got = h.download()
    rax, ray = oracle.accel(m, x, y, vx, vy, 3, nthreads=NTHREADS)
    assert bits_equal(ax, rax) and bits_equal(ay, ray)
    want = oracle.step(m, x, y, vx, vy, 3, 1, nthreads=NTHREADS)
    for a, b in zip(got, want):
        assert bits_equal(a, b)


def test_strict_200_bodies_with_coincident_pairs():
    """The reference's MAX_OBJECT-sized case on the cooperative kernel, with exact
    duplicates (r == 0 skip) and a massless body."""
    n = 200
    m, x, y, vx, vy = g.plummer(n, seed=21)
    x[150], y[150], vx[150], vy[150] = x[3], y[3], vx[3], vy[3]
    vx[3] = vy[3] = vx[150] = vy[150] = 0.0
    m[77] = 0.0
    with g.GravityCuda(n) as h:
        h.upload(m, x, y, vx, vy)
        h.step(60, 1001)
        got = h.download()
    want = oracle.step(m, x, y, vx, vy, 60, 1001)
    for a, b in zip(got, want):
        assert bits_equal(a, b)


# ---------------------------------------------------------------------------
# TILED: tolerance parity
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("name", ok_fixture_names())
def test_tiled_accel_on_scenarios(name):
    fx = load_fixture(name)
    m, x, y, vx, vy = state0(fx)
    with g.GravityCuda(fx["n"], kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(fx["dt"])
    rax, ray = oracle.accel(m, x, y, vx, vy, fx["dt"])
    assert rel_err_rows(ax, ay, rax, ray).max() <= ACCEL_TOL


def assert_trajectory_parity(x, y, vx, vy, cp, label):
    """PER BODY |dp| / |p| and |dv| / |v| <= 1e-9.  A body sitting
    at (almost) the origin of the coordinates, or (almost) at rest, has no scale of its own: for
    |p| < 1e-3 of the largest |p| in the system (resp. |v|) the system's scale stands in."""
    pmag, vmag = np.hypot(cp["x_f"], cp["y_f"]), np.hypot(cp["vx_f"], cp["vy_f"])
    dp, dv = np.hypot(x - cp["x_f"], y - cp["y_f"]), np.hypot(vx - cp["vx_f"], vy - cp["vy_f"])
    ep = dp / np.maximum(pmag, 1e-3 * pmag.max())
    ev = dv / np.maximum(vmag, 1e-3 * vmag.max())
    assert ep.max() <= TRAJ_TOL, "%s: per-body position error %.3g (body %d)" % (label, ep.max(), int(ep.argmax()))
    assert ev.max() <= TRAJ_TOL, "%s: per-body velocity error %.3g (body %d)" % (label, ev.max(), int(ev.argmax()))
    return float(ep.max()), float(ev.max())


NON_CHAOTIC = ["ref_two_body_circular", "ref_two_body_elliptical", "ref_solar_system",
               "ref_solar_system_and_rogue_star", "ref_gravity_boost"]


@pytest.mark.parametrize("name", NON_CHAOTIC)
def test_tiled_trajectory_1e4_steps(name):
    """Trajectories <= 1e-9 relative after 10^4 steps; three_body is excluded
    because the reference itself calls it chaotic (sim_gravity_help.h:63-65)."""
    fx = load_fixture(name)
    cp = [c for c in fx["checkpoints"] if c["steps"] == 10000][0]
    with g.GravityCuda(fx["n"], kernel=g.KERNEL_TILED) as h:
        h.upload(*state0(fx))
        h.step(fx["dt"], 10000)
        x, y, vx, vy = h.download()
    assert_trajectory_parity(x, y, vx, vy, cp, name)
    # total energy drift after the 10^4 steps matches the reference's drift
    m = fx["mass_f"]
    e0 = sum(oracle.energy(*state0(fx)))
    drift_ref = (sum(oracle.energy(m, cp["x_f"], cp["y_f"], cp["vx_f"], cp["vy_f"])) - e0) / abs(e0)
    drift_gpu = (sum(oracle.energy(m, x, y, vx, vy)) - e0) / abs(e0)
    assert abs(drift_gpu - drift_ref) <= TRAJ_TOL


def test_tiled_rogue_star_1e5_steps_and_energy_drift():
    """BASELINE.json configs[1]: 10^5 steps, checked every 10^4th step against
    the reference, energy-drift curve within the same tolerance."""
    fx = load_fixture("ref_solar_system_and_rogue_star")
    m = fx["mass_f"]
    cps = [c for c in fx["checkpoints"] if c["steps"] % 10000 == 0 and c["steps"] > 0]
    e0 = sum(oracle.energy(*state0(fx)))
    with g.GravityCuda(fx["n"], kernel=g.KERNEL_TILED) as h:
        h.upload(*state0(fx))
        done = 0
        for cp in cps:
            h.step(fx["dt"], cp["steps"] - done)
            done = cp["steps"]
            x, y, vx, vy = h.download()
            assert_trajectory_parity(x, y, vx, vy, cp, "rogue star step %d" % done)
            drift_ref = (sum(oracle.energy(m, cp["x_f"], cp["y_f"], cp["vx_f"], cp["vy_f"])) - e0) / abs(e0)
            ke, pe = h.energy()
            drift_gpu = (ke + pe - e0) / abs(e0)
            assert abs(drift_gpu - drift_ref) <= TRAJ_TOL, done


@pytest.mark.parametrize("n", [1025, 1300, 4096, 4097, 20000])
def test_tiled_accel_plummer_ragged_sizes(n):
    m, x, y, vx, vy = g.plummer(n, seed=1)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(1)
    rax, ray = oracle.accel(m, x, y, vx, vy, 1, nthreads=NTHREADS)
    assert_accel_parity(ax, ay, rax, ray, (m, x, y, vx, vy), 1)


@pytest.mark.parametrize("threads,bpt,cps", [(128, 1, 4), (128, 2, 3), (128, 4, 2), (256, 1, 2), (256, 2, 2), (256, 4, 1)])
def test_tiled_kernel_shapes_agree(threads, bpt, cps):
    n = 6000
    m, x, y, vx, vy = g.plummer(n, seed=5)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED, threads=threads, bodies_per_thread=bpt, ctas_per_sm=cps) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(3)
        h.step(3, 2)
        got = h.download()
    rax, ray = oracle.accel(m, x, y, vx, vy, 3, nthreads=NTHREADS)
    assert_accel_parity(ax, ay, rax, ray, (m, x, y, vx, vy), 3)
    want = oracle.step(m, x, y, vx, vy, 3, 2, nthreads=NTHREADS)
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-11, atol=0)


@pytest.mark.parametrize("fold", [0, 1])
def test_mass_fold_paths_agree_on_uniform_and_mixed_masses(fold):
    """mass_fold sums equal-mass j-tiles without the mass (12 FP64 ops per pair)
    and scales once per tile; both paths must meet the same bar, on an
    equal-mass system, on random masses and on a mixture of the two."""
    n = 9000
    m, x, y, vx, vy = g.plummer(n, seed=12)
    rng = np.random.default_rng(5)
    mixed = m.copy()
    mixed[4000:] *= rng.uniform(0.1, 10.0, n - 4000)       # tiles 0..14 uniform, 15 mixed, rest random
    mixed[123] = 0.0                                        # a massless tracer inside a uniform tile
    for masses in (m, mixed):
        with g.GravityCuda(n, kernel=g.KERNEL_TILED, mass_fold=fold) as h:
            h.upload(masses, x, y, vx, vy)
            ax, ay = h.accel(4)
            h.step(4, 2)
            got = h.download()
        rax, ray = oracle.accel(masses, x, y, vx, vy, 4, nthreads=NTHREADS)
        assert_accel_parity(ax, ay, rax, ray, (masses, x, y, vx, vy), 4)
        want = oracle.step(masses, x, y, vx, vy, 4, 2, nthreads=NTHREADS)
        for a, b in zip(got, want):
            assert np.allclose(a, b, rtol=1e-11, atol=0)


def test_tiled_accel_config3_full_65536():
    """BASELINE.json configs[2]: N = 65,536 on one GPU, accelerations against the
    oracle on the full set (4.3e9 pairs on the host)."""
    n = 65536
    m, x, y, vx, vy = g.plummer(n, seed=1)
    with g.GravityCuda(n) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(1)
    rax, ray = oracle.accel(m, x, y, vx, vy, 1, nthreads=NTHREADS)
    err = assert_accel_parity(ax, ay, rax, ray, (m, x, y, vx, vy), 1)
    assert np.median(err) <= 1e-14
    # and no farther from a long-double evaluation than the reference itself is
    rows = np.arange(0, n, 64, dtype=np.int64)
    tax, tay = oracle.accel_ld(m, x, y, vx, vy, 1, rows=rows, nthreads=NTHREADS)
    e_gpu = rel_err_rows(ax[rows], ay[rows], tax, tay)
    e_ref = rel_err_rows(rax[rows], ray[rows], tax, tay)
    assert np.median(e_gpu) <= 2 * np.median(e_ref) + 1e-16


def test_tiled_accel_config4_262144_row_sample():
    """BASELINE.json configs[3] size: N = 262,144, parity on a 4,096-row subsample."""
    n = 262144
    m, x, y, vx, vy = g.plummer(n, seed=1)
    with g.GravityCuda(n) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(1)
    rows = np.arange(0, n, n // 4096, dtype=np.int64)
    rax, ray = oracle.accel(m, x, y, vx, vy, 1, rows=rows, nthreads=NTHREADS)
    pure = rel_err_rows(ax[rows], ay[rows], rax, ray)
    floor = 0.1 * np.median(np.hypot(rax, ray))
    assert (np.hypot(ax[rows] - rax, ay[rows] - ray) / np.maximum(np.hypot(rax, ray), floor)).max() <= ACCEL_TOL
    assert (pure > ACCEL_TOL).sum() <= 1


def test_tiled_coincident_bodies_take_the_guarded_path():
    """Exactly coincident pairs are skipped (sim_gravity.c:1199-1201); the fast
    loop poisons those bodies with NaN and the commit kernel repairs them."""
    n = 5000
    m, x, y, vx, vy = g.plummer(n, seed=9)
    for a, b in ((10, 4000), (11, 12), (2999, 3000)):
        x[b], y[b] = x[a], y[a]
        vx[a] = vy[a] = 0.0        # a's half-drifted position is then exactly b's position
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(2)
        h.step(2, 1)
        got = h.download()
    assert np.isfinite(ax).all() and np.isfinite(ay).all()
    rax, ray = oracle.accel(m, x, y, vx, vy, 2, nthreads=NTHREADS)
    assert_accel_parity(ax, ay, rax, ray, (m, x, y, vx, vy), 2)
    want = oracle.step(m, x, y, vx, vy, 2, 1, nthreads=NTHREADS)
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-11, atol=0)


def test_self_term_is_skipped_by_index_not_by_distance():
    """Only the i-body is half-drifted (sim_gravity.c:1186-1197), so the self
    distance is |V*DT/2| != 0; a kernel that relied on r == 0 would add a huge
    bogus self-force.  Big DT and big velocities make that visible."""
    n = 2048
    m, x, y, vx, vy = g.plummer(n, seed=2)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx * 50, vy * 50)
        ax, ay = h.accel(86400)
    rax, ray = oracle.accel(m, x, y, vx * 50, vy * 50, 86400, nthreads=NTHREADS)
    assert_accel_parity(ax, ay, rax, ray, (m, x, y, vx * 50, vy * 50), 86400)


def test_bodies_at_absurd_distances_contribute_what_the_reference_computes():
    """pair_weight squares the reciprocal-square-root seed: for r between ~7e153 and 1.3e154 m that
    square is subnormal and the series correction meaningless -- but m / r^3 is then below the
    smallest double for any mass under ~1e138 kg, and the product underflows to the same exact 0
    the reference's m / (r*r*r) = m / inf gives.  Bodies parked there (the padding sits at 1e150),
    and at every decade up to overflow, must leave the other accelerations untouched."""
    n = 3000
    m, x, y, vx, vy = g.plummer(n, seed=14)
    far = [1e150, 3e152, 5e153, 9e153, 1.2e154, 1.3e154, 1e155, 1e200, 1e300]
    for k, r in enumerate(far):
        x[100 + k], y[100 + k] = r, -0.5 * r
        vx[100 + k] = vy[100 + k] = 0.0
    m[100] = 1e35                                            # a huge but finite mass far away
    with g.GravityCuda(n, kernel=g.KERNEL_TILED, mass_fold=0) as h:
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(5)
    rax, ray = oracle.accel(m, x, y, vx, vy, 5, nthreads=NTHREADS)
    assert np.isfinite(ax).all() and np.isfinite(ay).all()
    near = np.ones(n, bool)
    near[100:100 + len(far)] = False
    assert rel_err_rows(ax[near], ay[near], rax[near], ray[near]).max() <= ACCEL_TOL
    # the far bodies themselves: every pair distance is ~r, their acceleration underflows or is tiny
    assert np.allclose(ax[~near], rax[~near], rtol=1e-12, atol=1e-300) and np.allclose(ay[~near], ray[~near], rtol=1e-12, atol=1e-300)


def test_upload_without_masses_keeps_them():
    n = 5000
    m, x, y, vx, vy = g.plummer(n, seed=15)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        with pytest.raises(g.GravityError, match="mass"):
            h.upload(None, x, y, vx, vy)                     # nothing to keep yet
        h.upload(m, x, y, vx, vy)
        h.step(2, 3)
        a = h.download()
        h.upload(None, x, y, vx, vy)                         # same masses, state rewound
        h.step(2, 3)
        b = h.download()
        h.set_option("download_scope", 1)                    # one GPU: the shard is everything
        c = h.download()
    for p, q, r in zip(a, b, c):
        assert bits_equal(p, q) and bits_equal(p, r)


def test_step_host_is_upload_step_download_in_one_call():
    """gravity_cuda_step_host: pageable arrays (staging copies) and page-locked ones (read and
    written in place by the device), masses optional after the first call, arrays updated in place."""
    import torch
    n = 6000
    m, x, y, vx, vy = g.plummer(n, seed=18)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
        h.upload(m, x, y, vx, vy)
        h.step(4, 3)
        want3 = h.download()
        h.step(4, 2)
        want5 = h.download()
    for pinned in (False, True):
        arrs = [a.copy() for a in (x, y, vx, vy)]
        keep = None
        if pinned:
            keep = [torch.from_numpy(a).pin_memory() for a in arrs]
            arrs = [t.numpy() for t in keep]
        with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
            h.step_host(m, *arrs, 4, 3)
            for a, b in zip(arrs, want3):
                assert bits_equal(a, b), pinned
            h.step_host(None, *arrs, 4, 2)                   # the masses stay on the device
            for a, b in zip(arrs, want5):
                assert bits_equal(a, b), pinned
            with pytest.raises(ValueError):
                h.step_host(m, arrs[0][::2], *arrs[1:], 4, 1)
    with g.GravityCuda(n) as h:
        with pytest.raises(g.GravityError, match="mass"):
            h.step_host(None, x.copy(), y.copy(), vx.copy(), vy.copy(), 1, 1)


def test_batches_and_single_steps_give_the_same_bits():
    """One launch for the whole batch (steps chained by per-i-block ready counters) against one
    launch per step: same plan, same summation order, identical bits; and the batch is repeatable."""
    n = 7000
    m, x, y, vx, vy = g.plummer(n, seed=16)
    runs = []
    for persistent, chunks in ((1, [37]), (1, [1] * 37), (0, [37]), (1, [5, 32])):
        with g.GravityCuda(n, kernel=g.KERNEL_TILED, persistent=persistent) as h:
            h.upload(m, x, y, vx, vy)
            for c in chunks:
                h.step(3, c)
            runs.append(h.download())
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert bits_equal(a, b)
    want = oracle.step(m, x, y, vx, vy, 3, 37, nthreads=NTHREADS)
    for a, b in zip(runs[0], want):
        assert np.allclose(a, b, rtol=1e-11, atol=0)


@pytest.mark.parametrize("n,threads,bpt,cps", [(1025, 0, 0, 0), (2049, 128, 1, 4), (4097, 128, 2, 3), (4097, 256, 1, 2),
                                               (9001, 256, 2, 2), (9001, 256, 4, 1), (20011, 0, 0, 0)])
def test_many_steps_in_one_launch_stay_bit_identical_to_one_launch_per_step(n, threads, bpt, cps):
    """Stress of the in-kernel data flow (arrival / ready counters, TMA fetches of positions other
    CTAs have just committed, slices, the two ping-pong buffers): 400 steps in ONE launch against
    400 launches, small systems where a step is tens of microseconds and CTAs run far apart.
    Any ordering hole shows up as a differing bit."""
    m, x, y, vx, vy = g.plummer(n, seed=40 + bpt)
    opts = dict(threads=threads, bodies_per_thread=bpt, ctas_per_sm=cps) if threads else {}
    out = []
    for persistent in (1, 0):
        with g.GravityCuda(n, kernel=g.KERNEL_TILED, persistent=persistent, **opts) as h:
            h.upload(m, x, y, vx, vy)
            h.step(3600, 400)
            out.append(h.download())
    for a, b in zip(*out):
        assert bits_equal(a, b)
    assert np.isfinite(out[0][0]).all()


def test_energy_matches_oracle():
    n = 10000
    m, x, y, vx, vy = g.plummer(n, seed=4)
    with g.GravityCuda(n) as h:
        h.upload(m, x, y, vx, vy)
        ke, pe = h.energy()
    rke, rpe = oracle.energy(m, x, y, vx, vy, nthreads=NTHREADS)
    assert abs(ke - rke) / rke < 1e-13
    assert abs(pe - rpe) / abs(rpe) < 1e-12


def test_step_objects_aos_adapter():
    """The reference keeps bodies as individually allocated object_t records
    (X@32, Y@40, VX@48, VY@56, MASS@64; sim_gravity.c:83-105)."""
    fx = load_fixture("ref_gravity_boost")
    m, x, y, vx, vy = state0(fx)
    n = fx["n"]
    recs = [ctypes.create_string_buffer(256) for _ in range(n)]
    for i, r in enumerate(recs):
        for off, v in ((32, x[i]), (40, y[i]), (48, vx[i]), (56, vy[i]), (64, m[i])):
            ctypes.c_double.from_buffer(r, off).value = v
    with g.GravityCuda(n) as h:
        h.step_objects([ctypes.addressof(r) for r in recs], 32, 40, 48, 56, 64, fx["dt"], 100)
    cp = [c for c in fx["checkpoints"] if c["steps"] == 100][0]
    got = np.array([[ctypes.c_double.from_buffer(r, off).value for off in (32, 40, 48, 56)] for r in recs])
    assert bits_equal(got[:, 0], cp["x_f"]) and bits_equal(got[:, 3], cp["vy_f"])


def test_error_behaviour():
    with pytest.raises(g.GravityError):
        g.GravityCuda(0)
    with g.GravityCuda(4) as h:
        with pytest.raises(g.GravityError, match="before upload"):
            h.step(1, 1)
        with pytest.raises(g.GravityError, match="unknown option"):
            h.set_option("nonsense", 1)


# ---------------------------------------------------------------------------
# Size-independent properties at BASELINE.json's full size
# ---------------------------------------------------------------------------

def test_full_size_1m_properties_and_row_sample():
    n = 1 << 20
    m, x, y, vx, vy = g.plummer(n, seed=1)
    zero = np.zeros(n)
    with g.GravityCuda(n) as h:
        # with V = 0 nothing is drifted, pair forces are antisymmetric and the
        # total force must vanish to rounding
        h.upload(m, x, y, zero, zero)
        ax, ay = h.accel(1)
        assert np.isfinite(ax).all() and np.isfinite(ay).all()
        fsum = np.hypot(np.sum(m * ax), np.sum(m * ay))
        assert fsum / np.sum(m * np.hypot(ax, ay)) < 1e-9
        # row sample against the oracle (1,024 rows x 1M bodies on the host)
        h.upload(m, x, y, vx, vy)
        ax, ay = h.accel(1)
        rows = np.arange(0, n, 1024, dtype=np.int64)
        rax, ray = oracle.accel(m, x, y, vx, vy, 1, rows=rows, nthreads=NTHREADS)
        pure = rel_err_rows(ax[rows], ay[rows], rax, ray)
        floor = 0.1 * np.median(np.hypot(rax, ray))
        assert (np.hypot(ax[rows] - rax, ay[rows] - ray) / np.maximum(np.hypot(rax, ray), floor)).max() <= ACCEL_TOL
        assert (pure > ACCEL_TOL).sum() <= 1 and np.median(pure) <= 1e-13   # the reference sum itself carries ~sqrt(N) eps
        # one step moves every body by V*DT + A*DT^2/2 (sim_gravity.c:1213-1219)
        h.step(1, 1)
        nx, ny, nvx, nvy = h.download()
        assert np.allclose(nx[rows], x[rows] + (vx[rows] + 0.5 * rax), rtol=1e-13, atol=0)
        assert np.allclose(nvy[rows], vy[rows] + ray, rtol=1e-12, atol=0)


# ---------------------------------------------------------------------------
# Multi-GPU (single-process handle; skipped on a one-GPU box)
# ---------------------------------------------------------------------------

@pytest.mark.multigpu
@pytest.mark.parametrize("exchange", [0, 1])
def test_two_gpus_agree_with_one(gpu_count, exchange):
    """exchange=0: NCCL all-gather; exchange=1: the step kernel stores into the
    peer's position buffer over NVLink and publishes a step flag."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    n = 30000
    m, x, y, vx, vy = g.plummer(n, seed=6)
    with g.GravityCuda(n, ngpus=1, kernel=g.KERNEL_TILED) as h1:
        h1.upload(m, x, y, vx, vy)
        h1.step(5, 4)
        one = h1.download()
    with g.GravityCuda(n, ngpus=2, kernel=g.KERNEL_TILED, exchange=exchange) as h2:
        h2.upload(m, x, y, vx, vy)
        h2.step(5, 3)
        h2.accel(5)                      # an evaluation that must not disturb the exchange
        h2.step(5, 1)
        ke, pe = h2.energy()
        two = h2.download()
        assert np.isfinite(ke + pe)
    for a, b in zip(one, two):
        assert np.allclose(a, b, rtol=1e-12, atol=0)
    want = oracle.step(m, x, y, vx, vy, 5, 4, nthreads=NTHREADS)
    for a, b in zip(two, want):
        assert np.allclose(a, b, rtol=1e-11, atol=0)


@pytest.mark.multigpu
@pytest.mark.parametrize("n", [3000, 20000])
def test_two_gpus_both_exchanges_give_the_same_bits_over_many_steps(gpu_count, n):
    """300 steps in one launch per GPU, chained by peer flags over NVLink, against 300 launches with an
    NCCL all-gather after each: same plan, same summation order, so every bit must agree -- a stale or
    torn remote position would not."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    m, x, y, vx, vy = g.plummer(n, seed=51)
    out = []
    for exchange in (1, 0):
        with g.GravityCuda(n, ngpus=2, kernel=g.KERNEL_TILED, exchange=exchange) as h:
            h.upload(m, x, y, vx, vy)
            h.step(3600, 300)
            out.append(h.download())
    for a, b in zip(*out):
        assert bits_equal(a, b)


@pytest.mark.multigpu
@pytest.mark.parametrize("exchange", [0, 1])
def test_two_gpus_strict_is_still_bit_exact(gpu_count, exchange):
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    n = 1500
    m, x, y, vx, vy = g.plummer(n, seed=8)
    with g.GravityCuda(n, ngpus=2, kernel=g.KERNEL_STRICT, exchange=exchange) as h:
        h.upload(m, x, y, vx, vy)
        h.step(9, 3)
        got = h.download()
    want = oracle.step(m, x, y, vx, vy, 9, 3, nthreads=NTHREADS)
    for a, b in zip(got, want):
        assert bits_equal(a, b)


# ---------------------------------------------------------------------------
# The C host: scenario file -> gravity_headless -> state, PATH sampling
# ---------------------------------------------------------------------------

def _reference_path_samples(steps, dt):
    """Replay of the commit block (sim_gravity.c:1229-1260): how many PATH
    samples the reference takes in `steps` steps from sim_time 0."""
    sim_time, last_day, tail = 0, 0, 0
    for _ in range(steps):
        day = sim_time // 86400
        if day != last_day:
            tail += 1
        last_day = day
        sim_time += dt
    return tail


@pytest.mark.parametrize("name", [n for n in ok_fixture_names("own_")])
def test_headless_c_driver_matches_reference_bits(name):
    import json
    import os
    import subprocess
    from conftest import ROOT, scenario_path
    fx = load_fixture(name)
    path = scenario_path(fx)
    steps = [c["steps"] for c in fx["checkpoints"]]
    exe = os.path.join(ROOT, "proj_entropy_b200", "gravity_headless")
    p = subprocess.run([exe, "--scenario", path, "--steps", str(steps[-1]),
                        "--checkpoint", ",".join(str(s) for s in steps)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    states, summary = lines[:-1], lines[-1]
    assert [s["steps"] for s in states] == steps
    for s, cp in zip(states, fx["checkpoints"]):
        assert s["sim_time"] == cp["sim_time"] and s["dt"] == fx["dt"] and s["n"] == fx["n"]
        for key in ("x", "y", "vx", "vy"):
            got = [float.fromhex(b[key]) for b in s["bodies"]]
            assert bits_equal(got, cp[key + "_f"]), "%s step %d %s" % (name, s["steps"], key)
    assert summary["path_samples"] == _reference_path_samples(steps[-1], fx["dt"]) == fx["checkpoints"][-1]["path_tail"]
    assert summary["path_hash"] == fx["checkpoints"][-1]["path_hash"]     # the reference's own PATH rings


def test_upload_download_round_trip_bits():
    n = 70001
    m, x, y, vx, vy = g.plummer(n, seed=11)
    with g.GravityCuda(n) as h:
        h.upload(m, x, y, vx, vy)
        got = h.download()
    for a, b in zip(got, (x, y, vx, vy)):
        assert bits_equal(a, b)


@pytest.mark.multigpu
def test_peer_exchange_with_an_empty_shard(gpu_count):
    """N smaller than one shard: the second GPU owns no body but still has to
    publish its step flag, or the first would wait for ever."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    n = 200
    m, x, y, vx, vy = g.plummer(n, seed=13)
    with g.GravityCuda(n, ngpus=2, kernel=g.KERNEL_TILED, exchange=1) as h:
        h.upload(m, x, y, vx, vy)
        h.step(3, 5)
        got = h.download()
    want = oracle.step(m, x, y, vx, vy, 3, 5)
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-11, atol=0)


def test_headless_checkpoint_resume_is_bit_exact(tmp_path):
    """Stop after 3 steps, save a snapshot, resume in a new process, run 3 more:
    identical to 6 steps straight (tiled kernel, 5,000 bodies)."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "proj_entropy_b200", "gravity_headless")
    snap = str(tmp_path / "mid.snap")

    def run(*args):
        p = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert p.returncode == 0, p.stderr
        return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]

    straight = run("--plummer", "5000", "--dt", "30000", "--steps", "6", "--checkpoint", "6", "--no-path")[0]
    run("--plummer", "5000", "--dt", "30000", "--steps", "3", "--no-path", "--save", snap)
    resumed = run("--resume", snap, "--steps", "3", "--checkpoint", "3", "--no-path")[0]
    assert resumed["sim_time"] == straight["sim_time"] == 6 * 30000
    for key in ("x", "y", "vx", "vy"):
        assert [b[key] for b in resumed["bodies"]] == [b[key] for b in straight["bodies"]]
