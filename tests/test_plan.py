"""The launch planner of the tiled kernel (host logic, no GPU): every
(i-block, j-tile) unit is computed exactly once, the CTAs' shares differ by at
most one unit, own-shard tiles come first in every CTA, and the partial rows of
an i-block are contiguous and in CTA order (what makes the sums deterministic)."""
import numpy as np
import pytest

import proj_entropy_b200 as g
from proj_entropy_b200.dist import shard_bounds

CASES = [(1025, 1, 0), (4097, 1, 0), (65536, 1, 0), (1 << 20, 1, 0), (65536, 8, 3), (65536, 8, 0), (65536, 8, 7),
         (262144, 4, 2), (1 << 20, 8, 5), (30000, 2, 1), (200, 2, 0), (5000, 8, 7), (16384, 2, 1), (1000003, 8, 7)]


@pytest.mark.parametrize("n,nranks,rank", CASES)
def test_plan_covers_every_unit_once_and_is_balanced(n, nranks, rank):
    info, segs = g.plan_probe(n, nranks=nranks, rank=rank)
    b, e = shard_bounds(n, nranks, rank)
    assert info["n_local"] == e - b
    ib = info["threads"] * info["bodies_per_thread"]
    assert info["n_iblock"] == -(-(e - b) // ib)
    assert info["n_tile"] == -(-n // 256)
    if e == b:
        assert info["n_cta"] == 0 and len(segs) == 0
        return
    cover = np.zeros((info["n_iblock"], info["n_tile"]), dtype=np.int32)
    per_cta = np.zeros(info["n_cta"], dtype=np.int64)
    for cta, iblock, t0, t1, row in segs:
        assert 0 <= t0 < t1 <= info["n_tile"]
        cover[iblock, t0:t1] += 1
        per_cta[cta] += t1 - t0
    assert (cover == 1).all()
    assert per_cta.max() - per_cta.min() <= 1
    assert info["n_cta"] <= 148 * info["ctas_per_sm"]
    assert sorted(segs[:, 4]) == list(range(len(segs)))          # one partial row per segment


@pytest.mark.parametrize("n,nranks,rank", [c for c in CASES if c[1] > 1])
def test_own_shard_tiles_come_first_in_every_cta(n, nranks, rank):
    info, segs = g.plan_probe(n, nranks=nranks, rank=rank)
    lo, hi = info["local_tile_begin"], info["local_tile_end"]
    b, e = shard_bounds(n, nranks, rank)
    if e > b:
        assert lo == b // 256 and hi >= -(-e // 256)
    for cta in range(info["n_cta"]):
        mine = segs[segs[:, 0] == cta]
        local = [(lo <= t0 and t1 <= hi) for _, _, t0, t1, _ in mine]
        assert all(lo <= t0 and t1 <= hi or t1 <= lo or t0 >= hi for _, _, t0, t1, _ in mine)   # never straddles
        first_remote = local.index(False) if False in local else len(local)
        assert not any(local[first_remote:]), "a local segment after a remote one in CTA %d" % cta


@pytest.mark.parametrize("n,nranks,rank", CASES)
def test_partial_rows_of_an_iblock_are_contiguous_and_in_cta_order(n, nranks, rank):
    info, segs = g.plan_probe(n, nranks=nranks, rank=rank)
    if len(segs) == 0:
        return
    by_row = segs[np.argsort(segs[:, 4])]
    assert (np.diff(by_row[:, 1]) >= 0).all()                  # rows grouped by i-block
    for ib in range(info["n_iblock"]):
        ctas = by_row[by_row[:, 1] == ib][:, 0]
        assert (np.diff(ctas) >= 0).all()                      # and ordered by CTA within it


def test_shape_follows_problem_size():
    assert g.plan_probe(1 << 20)[0]["threads"] * g.plan_probe(1 << 20)[0]["bodies_per_thread"] == 1024
    small = g.plan_probe(2048)[0]
    assert small["threads"] * small["bodies_per_thread"] < 1024    # more, smaller units to fill 148 SMs
    forced = g.plan_probe(2048, threads=256, bodies_per_thread=2)[0]
    assert (forced["threads"], forced["bodies_per_thread"], forced["ctas_per_sm"]) == (256, 2, 2)
    with pytest.raises(g.GravityError):
        g.plan_probe(0)


def test_plan_of_a_very_large_system_is_cheap_and_exact():
    # 2^26 bodies on 8 GPUs: 1.7e10 work units; the plan is O(CTAs + i-blocks)
    n = 1 << 26
    info, segs = g.plan_probe(n, nranks=8, rank=6)
    assert info["n_tile"] == n // 256 and info["n_iblock"] == n // 8 // 1024 and info["n_cta"] == 148
    per_cta = np.zeros(148, dtype=np.int64)
    units = 0
    for cta, iblock, t0, t1, row in segs.tolist():      # python ints: 2^31 units overflow int32
        per_cta[cta] += t1 - t0
        units += t1 - t0
    assert units == info["n_iblock"] * info["n_tile"] == 1 << 31
    assert per_cta.max() - per_cta.min() <= 1
