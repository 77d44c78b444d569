#!/usr/bin/env python3
"""Developer probe under torchrun: per-rank step time and per-launch durations of
the hot kernel, to see where multi-GPU runs lose time."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from proj_entropy_b200.dist import Ranks

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--sync-each-step", type=int, default=0)
ap.add_argument("--exchange", type=int, default=0)
args = ap.parse_args()
r = Ranks("nccl")
n = args.bodies
state = g.plummer(n, seed=1)
nid = r.broadcast_bytes(g.nccl_unique_id() if r.rank == 0 else None, 128) if r.world > 1 else None
h = g.GravityCuda(n, rank=r.rank, nranks=r.world, device=r.local_rank, nccl_id=nid, kernel=g.KERNEL_TILED,
                  mass_fold=0)
if r.world > 1 and args.exchange == 1:
    h.p2p_connect(r.all_gather_bytes(h.p2p_export()))
    h.set_option("exchange", 1)
h.upload(*state)
h.step(1, 3)
r.barrier()
h.timer_start()
for _ in range(args.steps):
    h.step_async(1, 1)
    if args.sync_each_step:
        h.sync()
ms = h.timer_stop()
times = h.kernel_times()
smi = os.popen("nvidia-smi -i %d --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader" % r.local_rank).read().strip()
for k in range(r.world):
    r.barrier()
    if k == r.rank:
        print(json.dumps({"rank": r.rank, "exchange": args.exchange, "ms_per_step": ms / args.steps,
                          "launch_ms": [round(float(t), 3) for t in times[:12]], "smi_after": smi}), flush=True)
h.close()
r.close()
