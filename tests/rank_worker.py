"""Run under torchrun on GPUs (one rank per GPU) by tests/test_gpu_rank_mode.py:
a rank-mode handle per process, both exchange mechanisms, checked on rank 0
against the ORACLE (the CPU restatement of the reference loop), never against
this library's own single-GPU path.

argv: out_dir exchange n [mode]     mode = parity (default) | timeout
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from proj_entropy_b200.dist import Ranks, shard_bounds

out_dir, exchange, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "parity"
r = Ranks("nccl")
m, x, y, vx, vy = g.plummer(n, seed=17)
nid = r.broadcast_bytes(g.nccl_unique_id() if r.rank == 0 else None, 128) if r.world > 1 else None
h = g.GravityCuda(n, rank=r.rank, nranks=r.world, device=r.local_rank, nccl_id=nid, kernel=g.KERNEL_TILED)
if exchange and r.world > 1:
    h.p2p_connect(r.all_gather_bytes(h.p2p_export()))
    h.set_option("exchange", 1)
res = {"rank": r.rank, "world": r.world, "mode": mode}

if mode == "timeout":
    # rank 0 steps twice, the others never step: the second step needs the peers' first
    # step and must give up after the limit instead of hanging or committing stale data
    h.set_option("spin_timeout_ms", 300)
    h.upload(m, x, y, vx, vy)
    r.barrier()
    if r.rank == 0:
        t0 = time.perf_counter()
        try:
            h.step(3, 2)
            res["error"] = None
        except g.GravityError as e:
            res["error"] = str(e)
        res["seconds"] = time.perf_counter() - t0
        try:
            h.step(3, 1)
            res["second_error"] = None
        except g.GravityError as e:
            res["second_error"] = str(e)
    else:
        time.sleep(3.0)
    r.barrier()
    h.close()
    json.dump(res, open(os.path.join(out_dir, "rank%d.json" % r.rank), "w"))
    r.close()
    sys.exit(0)

h.upload(m, x, y, vx, vy)
h.step(3, 2)
ax, ay = h.accel(3)                       # gathered from every rank
h.step(3, 3)
ke, pe = h.energy()
got = h.download()                        # full arrays on every rank
# the caller's own shard only: no gather, the rest of the arrays is left alone
h.set_option("download_scope", 1)
mine = h.download()
lo, hi = shard_bounds(n, r.world, r.rank)
res["shard_download_ok"] = bool(all(np.array_equal(a[lo:hi], b[lo:hi]) for a, b in zip(mine, got)) and
                                all(not a[:lo].any() and not a[hi:].any() for a in mine))
# the one-call path on the shard: upload + 2 steps + download, in place, compared with stepping on
hx, hy, hvx, hvy = (a.copy() for a in got)
h.step_host(m, hx, hy, hvx, hvy, 3, 2)
h.step(3, 0)
h.upload(None, *got)
h.step(3, 2)
two = h.download()
res["step_host_ok"] = bool(all(np.array_equal(a[lo:hi], b[lo:hi]) for a, b in zip((hx, hy, hvx, hvy), two)))
h.set_option("download_scope", 0)
# a second upload must be safe while peers are connected; the masses stay (mass = None)
h.upload(None, x, y, vx, vy)
h.step(3, 5)
again = h.download()
h.close()
res["same_after_reupload"] = bool(all(np.array_equal(a, b) for a, b in zip(got, again)))
if r.rank == 0:
    from oracle import oracle_py as oracle
    nthreads = max(1, min((os.cpu_count() or 2) - 1, 16))
    mid = oracle.step(m, x, y, vx, vy, 3, 2, nthreads=nthreads)
    rax, ray = oracle.accel(m, *mid, 3, nthreads=nthreads)
    want = oracle.step(m, *mid, 3, 3, nthreads=nthreads)
    rke, rpe = oracle.energy(m, *want, nthreads=nthreads)
    res["accel_err"] = float((np.hypot(ax - rax, ay - ray) / np.hypot(rax, ray)).max())
    res["state_err"] = float(max((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max() for a, b in zip(got, want)))
    res["pos_err"] = float((np.hypot(got[0] - want[0], got[1] - want[1]) / np.hypot(want[0], want[1])).max())
    res["vel_err"] = float((np.hypot(got[2] - want[2], got[3] - want[3]) / np.hypot(want[2], want[3])).max())
    res["energy_err"] = float(abs((ke + pe) - (rke + rpe)) / abs(rke + rpe))
np.save(os.path.join(out_dir, "x_rank%d.npy" % r.rank), got[0])
np.save(os.path.join(out_dir, "vy_rank%d.npy" % r.rank), got[3])
json.dump(res, open(os.path.join(out_dir, "rank%d.json" % r.rank), "w"))
r.close()
