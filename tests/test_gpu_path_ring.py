"""the once-per-simulated-day PATH sampling of the
reference's commit block (sim_gravity.c:1230-1260) done on the device, inside the
kernels that run a whole batch of steps per launch, against the reference's own
PATH rings (golden path_tail / path_hash, dumped by the Tier-A harness) and
against positions stepped one by one."""
import json
import os
import subprocess

import numpy as np
import pytest

import proj_entropy_b200 as g
from conftest import ROOT, bits_equal, load_fixture, path_digest
from oracle import oracle_py as oracle

pytestmark = pytest.mark.gpu


def sample_steps(nsteps, dt, sim_time=0, last_day=0):
    """Replay of the day test: 1-based numbers of the steps whose commit takes a sample."""
    out = []
    for s in range(1, nsteps + 1):
        day = sim_time // 86400
        if day != last_day:
            out.append(s)
        last_day = day
        sim_time += dt
    return out, last_day


def state0(fx):
    cp = fx["checkpoints"][0]
    return fx["mass_f"], cp["x_f"], cp["y_f"], cp["vx_f"], cp["vy_f"]


@pytest.mark.parametrize("name,batches", [("own_dt_big", [1, 2, 7, 90, 900]), ("own_path_hourly", [1000]),
                                          ("own_solar_system_daily", [10, 90, 900, 9000]),
                                          ("ref_solar_system_and_rogue_star", [100000])])
def test_device_ring_equals_the_reference_path_rings(name, batches):
    """STRICT kernels (pair-parallel multi-step kernel): ring contents == the PATH rings of the
    unmodified reference, by count and digest at every checkpoint the batches end on."""
    fx = load_fixture(name)
    dt = fx["dt"]
    want = {c["steps"]: c for c in fx["checkpoints"]}
    with g.GravityCuda(fx["n"]) as h:
        h.upload(*state0(fx))
        h.path_enable(10240)
        sim_time, last_day, done, launches0 = 0, 0, 0, h.launch_count()
        for b in batches:
            last_day, tail = h.step_path(dt, b, sim_time, last_day)
            sim_time += dt * b
            done += b
            cp = want[done]
            assert tail == cp["path_tail"], (name, done)
            xs, ys = h.path_read(0, tail)
            assert path_digest(xs, ys) == cp["path_hash"], (name, done)
            x, y, vx, vy = h.download()
            assert bits_equal(x, cp["x_f"]) and bits_equal(vy, cp["vy_f"])
        # one launch per batch, whatever its length: no host round trip per simulated day
        assert h.launch_count() - launches0 <= 4 * len(batches)


@pytest.mark.parametrize("n,kernel,persistent", [(200, g.KERNEL_STRICT, 1), (1000, g.KERNEL_STRICT, 1),
                                                 (1500, g.KERNEL_STRICT, 1), (2000, g.KERNEL_TILED, 1),
                                                 (2000, g.KERNEL_TILED, 0), (9000, g.KERNEL_TILED, 1)])
def test_ring_rows_are_the_committed_positions(n, kernel, persistent):
    """Cooperative strict kernel, strict row kernel, tiled kernel (one launch per batch and one
    launch per step): the samples are exactly the positions of the sampled commits, the ring
    wraps like PATH[path_tail % MAX_PATH], and a batch may start mid-day."""
    dt, nsteps, cap = 30000, 40, 8
    m, x, y, vx, vy = g.plummer(n, seed=31)
    sim_time0, last_day0 = 50000, 0
    expect, last_day_end = sample_steps(nsteps, dt, sim_time0, last_day0)
    assert len(expect) > cap                                   # the ring wraps
    with g.GravityCuda(n, kernel=kernel, persistent=persistent) as ref:
        ref.upload(m, x, y, vx, vy)
        rows = {}
        for s in range(1, nsteps + 1):
            ref.step(dt, 1)
            if s in expect:
                rows[s] = ref.download()[:2]
        final = ref.download()
    with g.GravityCuda(n, kernel=kernel, persistent=persistent) as h:
        h.upload(m, x, y, vx, vy)
        h.path_enable(cap)
        ld, tail = h.step_path(dt, 15, sim_time0, last_day0)
        ld, tail = h.step_path(dt, nsteps - 15, sim_time0 + 15 * dt, ld)
        assert tail == len(expect) and ld == last_day_end
        xs, ys = h.path_read(tail - cap, cap)
        for k in range(cap):
            rx, ry = rows[expect[tail - cap + k]]
            assert bits_equal(xs[k], rx) and bits_equal(ys[k], ry), k
        with pytest.raises(g.GravityError, match="overwritten"):
            h.path_read(0, 1)
        with pytest.raises(g.GravityError, match="never taken"):
            h.path_read(tail, 1)
        got = h.download()
    for a, b in zip(got, final):
        assert bits_equal(a, b)
    if kernel == g.KERNEL_STRICT:
        want = oracle.step(m, x, y, vx, vy, dt, nsteps, nthreads=16)
        for a, b in zip(got, want):
            assert bits_equal(a, b)


def test_step_path_argument_checks():
    m, x, y, vx, vy = g.plummer(10, seed=1)
    with g.GravityCuda(10) as h:
        h.upload(m, x, y, vx, vy)
        with pytest.raises(g.GravityError, match="path_enable"):
            h.step_path(1, 1, 0, 0)
        h.path_enable(4)
        with pytest.raises(g.GravityError, match="dt >= 1"):
            h.step_path(0, 1, 0, 0)
        assert h.step_path(86400, 3, 0, 0) == (2, 2)
        h.path_enable(0)                                        # frees the ring
        with pytest.raises(g.GravityError, match="path_enable"):
            h.path_read(0, 1)


@pytest.mark.multigpu
@pytest.mark.parametrize("exchange", [0, 1])
def test_ring_on_two_gpus(gpu_count, exchange):
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    n, dt, nsteps = 5000, 43200, 9
    m, x, y, vx, vy = g.plummer(n, seed=33)
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as one:
        one.upload(m, x, y, vx, vy)
        one.path_enable(16)
        ld1, t1 = one.step_path(dt, nsteps, 0, 0)
        x1, y1 = one.path_read(0, t1)
    with g.GravityCuda(n, ngpus=2, kernel=g.KERNEL_TILED, exchange=exchange) as two:
        two.upload(m, x, y, vx, vy)
        two.path_enable(16)
        ld2, t2 = two.step_path(dt, nsteps, 0, 0)
        x2, y2 = two.path_read(0, t2)
    assert (ld1, t1) == (ld2, t2) == (4, 4)
    assert np.allclose(x1, x2, rtol=1e-12, atol=0) and np.allclose(y1, y2, rtol=1e-12, atol=0)


def test_headless_daily_steps_need_no_download_per_day():
    """`gravity_headless --scenario solar_system_daily --steps 10000` (DT = 86400 s: every commit
    is a PATH sample): O(1) ring downloads, and the rings equal the reference's bit for bit."""
    fx = load_fixture("own_solar_system_daily")
    exe = os.path.join(ROOT, "proj_entropy_b200", "gravity_headless")
    scen = os.path.join(ROOT, "tests", "golden", "scenarios", "solar_system_daily")
    p = subprocess.run([exe, "--scenario", scen, "--steps", "10000", "--checkpoint", "10000", "--path-batch", "4096"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    state, summary = lines[0], lines[-1]
    cp = fx["checkpoints"][-1]
    assert cp["steps"] == 10000 and state["sim_time"] == cp["sim_time"] == summary["sim_time"]
    assert summary["path_samples"] == cp["path_tail"] == 9999
    assert summary["path_hash"] == cp["path_hash"]
    assert summary["path_reads"] <= 3                                  # 10^4 samples through a 4,096-row ring
    assert summary["kernel_launches"] < 40                             # not one launch per step
    for key in ("x", "y", "vx", "vy"):
        assert bits_equal([float.fromhex(b[key]) for b in state["bodies"]], cp[key + "_f"]), key


def test_headless_resume_continues_the_path_rings(tmp_path):
    """2,000 daily steps straight, and 1,200 + a snapshot + 800 in a new process: the same PATH rings
    (count and digest -- the snapshot carries the samples taken so far), the same state, the same clock."""
    exe = os.path.join(ROOT, "proj_entropy_b200", "gravity_headless")
    scen = os.path.join(ROOT, "tests", "golden", "scenarios", "solar_system_daily")
    snap = str(tmp_path / "mid.snap")

    def run(*args):
        p = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert p.returncode == 0, p.stderr
        return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]

    straight = run("--scenario", scen, "--steps", "2000", "--checkpoint", "2000", "--path-batch", "512")
    first = run("--scenario", scen, "--steps", "1200", "--path-batch", "512", "--save", snap)
    resumed = run("--resume", snap, "--steps", "800", "--checkpoint", "800", "--path-batch", "300")
    assert first[-1]["path_samples"] == 1199
    a, b = straight[-1], resumed[-1]
    assert a["path_samples"] == b["path_samples"] == 1999 and a["path_hash"] == b["path_hash"]
    assert a["sim_time"] == b["sim_time"] and a["last_day"] == b["last_day"]
    for key in ("x", "y", "vx", "vy"):
        assert [q[key] for q in straight[0]["bodies"]] == [q[key] for q in resumed[0]["bodies"]]
