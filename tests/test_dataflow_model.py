"""A model of the step kernel's data flow, run on the REAL launch plans
(gravity_cuda_plan_probe, no GPU needed) under random schedules.

k_tiled_steps (csrc/gravity_kernels.cuh) runs every step of a batch in one
launch and replaces the reference's two spin barriers (sim_gravity.c:1140,
:1226) by counters: partial-row arrivals per i-block, slice commits, a ready
number per i-block, and the peers' step flags.  This test replays exactly that
protocol -- same segment lists, same waits, same slice split, same two ping-pong
position buffers -- with every CTA of every rank as a coroutine and a random
scheduler, and checks what the kernel's header comment claims:

  * no deadlock: whatever the interleaving, every CTA finishes every step;
  * no stale read: every i-body load and every j-tile fetch of step s sees
    positions of step s, for all bodies of the tile, local or remote;
  * no early overwrite: nobody stores step s+2 data into a buffer somebody is
    still reading for step s (two buffers suffice);
  * every body is committed exactly once per step.
"""
import random

import numpy as np
import pytest

import proj_entropy_b200 as g
from proj_entropy_b200.dist import shard_bounds

TILE = 256


class Rank:
    def __init__(self, n, nranks, rank, sms, threads, bpt, ctas):
        self.info, segs = g.plan_probe(n, nranks=nranks, rank=rank, num_sms=sms, threads=threads, bodies_per_thread=bpt,
                                       ctas_per_sm=ctas)
        self.rank, self.n = rank, n
        self.ib = self.info["threads"] * self.info["bodies_per_thread"]
        self.i0, hi = shard_bounds(n, nranks, rank)
        self.n_local = hi - self.i0
        self.n_iblock = self.info["n_iblock"]
        self.ctas = [[] for _ in range(self.info["n_cta"])]
        for cta, iblock, t0, t1, slot in segs.tolist():
            self.ctas[cta].append((iblock, t0, t1, slot))
        rows = np.zeros(self.n_iblock, dtype=np.int64)
        first = np.full(self.n_iblock, 1 << 30, dtype=np.int64)
        for cta in self.ctas:
            for iblock, _, _, slot in cta:
                rows[iblock] += 1
                first[iblock] = min(first[iblock], slot)
        self.rows, self.row0 = rows, first
        self.n_seg = int(rows.sum())
        self.local_tiles = (self.info["local_tile_begin"], self.info["local_tile_end"])
        npad = ((n + nranks * TILE - 1) // (nranks * TILE)) * nranks * TILE + nranks * TILE
        # what the device holds: the step number of every body's position in both buffers
        self.pos = np.zeros((2, npad), dtype=np.int64)
        self.pos[1, :] = -1                       # never written yet
        self.vel = np.zeros(max(1, self.n_local), dtype=np.int64)
        self.arrivals = np.zeros(self.n_iblock, dtype=np.int64)
        self.slices = np.zeros(self.n_iblock, dtype=np.int64)
        self.ready = np.zeros(self.n_iblock, dtype=np.int64)
        self.flags = np.zeros(nranks, dtype=np.int64)      # written by the peers
        self.xfer = 0
        self.committed = np.zeros(max(1, self.n_local), dtype=np.int64)


def cta_program(ranks, me, cta, nsteps, p2p=True, honor_ready=True):
    """One CTA of rank `me`: yields while it has to wait, mirrors k_tiled_steps.
    honor_ready=False drops the waits on the i-block ready numbers (to show that the model notices)."""
    R = ranks[me]
    segs = R.ctas[cta]
    others = [r for r in range(len(ranks)) if r != me]
    for s in range(nsteps):
        cur = s & 1
        peers_pending = p2p and len(ranks) > 1
        ready_lo, ready_hi = 0, -1
        for iblock, t0, t1, slot in segs:
            if s > 0 and honor_ready:
                while R.ready[iblock] < s:
                    yield
            lo = iblock * R.ib
            hi = min(R.n_local, lo + R.ib)
            assert (R.pos[cur, R.i0 + lo:R.i0 + hi] == s).all(), "stale i-bodies"
            assert (R.vel[lo:hi] == s).all(), "stale velocities"
            for t in range(t0, t1):
                if R.local_tiles[0] <= t < R.local_tiles[1]:
                    if s > 0 and honor_ready:
                        first = t * TILE - R.i0
                        b_lo, b_hi = first // R.ib, min(R.n_iblock - 1, (first + TILE - 1) // R.ib)
                        for b in range(b_lo, b_hi + 1):
                            if ready_lo <= b <= ready_hi:
                                continue
                            while R.ready[b] < s:
                                yield
                        if b_lo <= b_hi:
                            ready_lo, ready_hi = b_lo, b_hi
                elif peers_pending:
                    while any(R.flags[r] < s for r in others):
                        yield
                    peers_pending = False
                yield                              # the fetch itself takes time: others run meanwhile
                seen = R.pos[cur, t * TILE:min(R.n, (t + 1) * TILE)]
                assert (seen == s).all(), "tile %d of step %d holds %s" % (t, s, sorted(set(seen.tolist())))
            R.arrivals[iblock] += 1
        for iblock, t0, t1, slot in segs:
            while R.arrivals[iblock] < R.rows[iblock] * (s + 1):
                yield
            k, rows = slot - R.row0[iblock], R.rows[iblock]
            lo, hi = (k * R.ib) // rows, ((k + 1) * R.ib) // rows
            a, b = iblock * R.ib + lo, min(R.n_local, iblock * R.ib + hi)
            if b > a:
                yield                              # between reading the rows and storing the result
                assert (R.pos[cur ^ 1, R.i0 + a:R.i0 + b] <= s - 1).all(), "slice stored twice in one step"
                R.pos[cur ^ 1, R.i0 + a:R.i0 + b] = s + 1
                R.vel[a:b] = s + 1
                R.committed[a:b] += 1
                if p2p:
                    for r in others:               # stores into the peers' next buffer over NVLink
                        ranks[r].pos[cur ^ 1, R.i0 + a:R.i0 + b] = s + 1
            R.slices[iblock] += 1
            if R.slices[iblock] == rows * (s + 1):
                R.ready[iblock] = s + 1
        if p2p and len(ranks) > 1:
            yield                                  # the system-scope fence
            R.xfer += len(segs)
            if R.xfer == R.n_seg * (s + 1):
                for r in others:
                    ranks[r].flags[me] = s + 1


def run(n, nranks, sms, shape, nsteps, seed):
    threads, bpt, ctas = shape
    ranks = [Rank(n, nranks, r, sms, threads, bpt, ctas) for r in range(nranks)]
    rng = random.Random(seed)
    for R in ranks:
        if R.n_local == 0:                         # an empty shard publishes the whole batch at once (k_publish_only)
            for other in ranks:
                other.flags[R.rank] = nsteps
    progs = [cta_program(ranks, r, c, nsteps) for r in range(nranks) for c in range(len(ranks[r].ctas))]
    live = list(range(len(progs)))
    idle = 0
    while live:
        # a random CTA advances; CTAs that only spin change nothing, so count steps without any change
        k = rng.randrange(len(live))
        before = snapshot(ranks)
        try:
            next(progs[live[k]])
        except StopIteration:
            live.pop(k)
            idle = 0
            continue
        idle = 0 if snapshot(ranks) != before else idle + 1
        assert idle < 400 * len(progs) + 2000, "deadlock: %d CTAs spin and nothing changes" % len(live)
    for R in ranks:
        assert (R.committed[:R.n_local] == nsteps).all()
        assert (R.pos[nsteps & 1, :n] == nsteps).all(), "rank %d does not hold every body's final position" % R.rank
    return ranks


def snapshot(ranks):
    return tuple((int(R.arrivals.sum()), int(R.slices.sum()), int(R.ready.sum()), int(R.flags.sum()), R.xfer,
                  int(R.committed.sum())) for R in ranks)


CASES = [
    # n, ranks, SMs, (threads, bodies/thread, CTAs/SM)
    (1025, 1, 4, (128, 1, 4)), (3000, 1, 8, (128, 2, 3)), (3000, 1, 5, (256, 1, 2)), (5000, 1, 16, (256, 4, 1)),
    (2049, 1, 148, (0, 0, 0)), (6000, 1, 7, (256, 2, 2)), (777, 1, 3, (128, 1, 4)),
    (3000, 2, 4, (128, 1, 4)), (5000, 2, 6, (256, 2, 2)), (4100, 3, 5, (128, 2, 4)), (9000, 4, 4, (256, 4, 1)),
    (2600, 8, 2, (128, 1, 2)), (300, 2, 4, (128, 1, 4)),           # the last: rank 1 owns nothing
]


@pytest.mark.parametrize("n,nranks,sms,shape", CASES)
def test_data_flow_has_no_deadlock_and_no_stale_or_torn_read(n, nranks, sms, shape):
    for seed in range(3):
        ranks = run(n, nranks, sms, shape, nsteps=4, seed=seed)
        assert ranks is not None


def test_the_model_notices_a_missing_wait():
    """Sanity of the checker itself: without the i-block ready numbers a CTA that runs ahead reads
    positions of the wrong step, and the model says so."""
    n, nranks, sms, shape = 3000, 1, 8, (128, 2, 3)
    hit = 0
    for seed in range(6):
        ranks = [Rank(n, nranks, r, sms, *shape) for r in range(nranks)]
        rng = random.Random(seed)
        progs = [cta_program(ranks, 0, c, 3, honor_ready=False) for c in range(len(ranks[0].ctas))]
        live = list(range(len(progs)))
        try:
            steps = 0
            while live and steps < 200000:
                k = rng.randrange(len(live))
                steps += 1
                try:
                    next(progs[live[k]])
                except StopIteration:
                    live.pop(k)
        except AssertionError:
            hit += 1
    assert hit >= 1
