"""gravity_steps_until_day_change against a literal replay of the reference's
commit block (sim_gravity.c:1229-1260)."""
import itertools

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
