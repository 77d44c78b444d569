"""Run under torchrun on GPUs (one rank per GPU) by tests/test_gpu_rank_mode.py:
a rank-mode handle per process, both exchange mechanisms, checked against a
single-GPU run made by rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from proj_entropy_b200.dist import Ranks

out_dir, exchange, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
r = Ranks("nccl")
m, x, y, vx, vy = g.plummer(n, seed=17)
nid = r.broadcast_bytes(g.nccl_unique_id() if r.rank == 0 else None, 128) if r.world > 1 else None
h = g.GravityCuda(n, rank=r.rank, nranks=r.world, device=r.local_rank, nccl_id=nid, kernel=g.KERNEL_TILED)
if exchange and r.world > 1:
    h.p2p_connect(r.all_gather_bytes(h.p2p_export()))
    h.set_option("exchange", 1)
h.upload(m, x, y, vx, vy)
h.step(3, 2)
ax, ay = h.accel(3)                       # gathered from every rank
h.step(3, 3)
ke, pe = h.energy()
got = h.download()                        # full arrays on every rank
h.upload(m, x, y, vx, vy)                 # a second upload must be safe while peers are connected
h.step(3, 5)
again = h.download()
h.close()
res = {"rank": r.rank, "world": r.world, "same_after_reupload": all(np.array_equal(a, b) for a, b in zip(got, again))}
if r.rank == 0:
    with g.GravityCuda(n, kernel=g.KERNEL_TILED) as one:
        one.upload(m, x, y, vx, vy)
        one.step(3, 2)
        rax, ray = one.accel(3)
        one.step(3, 3)
        rke, rpe = one.energy()
        want = one.download()
    res["accel_err"] = float((np.hypot(ax - rax, ay - ray) / np.hypot(rax, ray)).max())
    res["state_err"] = float(max(np.abs(a - b).max() / np.abs(b).max() for a, b in zip(got, want)))
    res["energy_err"] = float(abs((ke + pe) - (rke + rpe)) / abs(rke + rpe))
np.save(os.path.join(out_dir, "x_rank%d.npy" % r.rank), got[0])
np.save(os.path.join(out_dir, "vy_rank%d.npy" % r.rank), got[3])
json.dump(res, open(os.path.join(out_dir, "rank%d.json" % r.rank), "w"))
r.close()
