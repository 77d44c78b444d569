"""gravity_steps_until_day_change against a literal replay of the reference's
commit block (sim_gravity.c:1229-1260)."""
import itertools

import pytest

import proj_entropy_b200 as g


def replay(sim_time, dt, last_day, max_steps=200000):
    """Number of commits until day_changed is first true."""
    for s in range(1, max_steps + 1):
        day = sim_time // 86400
        if day != last_day:
            return s
        sim_time += dt
    return None


def test_matches_replay():
    for dt, t0, last in itertools.product([1, 5, 7, 60, 3600, 86400, 100000], [0, 1, 86399, 86400, 172799, 250000],
                                           [0, 1, 2]):
        want = replay(t0, dt, last)
        if want is None:
            continue
        assert g.steps_until_day_change(t0, dt, last) == want, (dt, t0, last)


def test_zero_dt_never_samples():
    assert g.steps_until_day_change(0, 0, 0) == 2 ** 63 - 1


def test_device_clock_matches_a_replay_of_the_commit_block():
    """gravity_cuda_path_samples_in runs on the host the very clock_tick the step kernels carry
    (csrc/gravity_kernels.cuh): against a literal replay of sim_gravity.c:1232-1260 -- `int32_t day
    = sim_time/86400` before `sim_time += DT` -- for table and off-table DT, DT above a day, batches
    that start mid-day, and sim_time beyond the int32 range of the day number."""
    import ctypes
    import random
    import proj_entropy_b200 as g
    rng = random.Random(7)

    def replay(dt, nsteps, sim_time, last_day):
        tail = 0
        for _ in range(nsteps):
            day = ctypes.c_int32(sim_time // 86400).value
            if day != last_day:
                tail += 1
            last_day = day
            sim_time += dt
        return tail, last_day

    cases = [(1, 100000, 0, 0), (86400, 1000, 0, 0), (3600, 1000, 0, 0), (43200, 5, 0, 0), (30000, 40, 50000, 0),
             (200000, 50, 12345, 7), (86399, 30, 86399, 1), (86401, 30, 1, 0), (2147483647, 5, 0, 0),
             (86400, 10, 86400 * (2 ** 31 + 5), 3), (1, 5, 86400 * 2 ** 32 - 2, -1)]
    for _ in range(300):
        cases.append((rng.choice([1, 7, 60, 3600, 21600, 86400, rng.randint(1, 400000)]), rng.randint(0, 400),
                      rng.randint(0, 10 ** rng.randint(0, 15)), rng.randint(-3, 3)))
    for dt, nsteps, sim_time, last_day in cases:
        assert g.path_samples_in(dt, nsteps, sim_time, last_day) == replay(dt, nsteps, sim_time, last_day), (dt, nsteps, sim_time, last_day)
    for bad in ((0, 1, 0), (1, -1, 0), (1, 1, -5)):
        with pytest.raises(g.GravityError):
            g.path_samples_in(*bad)
