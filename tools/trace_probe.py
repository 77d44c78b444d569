#!/usr/bin/env python3
"""Developer probe: where does a step of the tiled kernel spend its time?  Runs a
batch with option "trace" = 1 and summarises the per-CTA globaltimer stamps
(gravity_cuda_trace_read): start skew, first-tile latency, pair phase, commit duty.

usage: python tools/trace_probe.py --bodies N [--shard-of P] [--steps K] [--persistent 0|1] [--opt k=v ...]
  --shard-of P   one GPU plans and runs like rank 0 of P ranks (option debug_shard_of): the
                 per-GPU work of an N-body run on P GPUs, without the exchange
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=65536)
ap.add_argument("--shard-of", type=int, default=1)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--persistent", type=int, default=1)
ap.add_argument("--mass-fold", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
args = ap.parse_args()

n = args.bodies
state = g.plummer(n, seed=1)
opts = {"kernel": g.KERNEL_TILED, "mass_fold": args.mass_fold, "trace": 1, "persistent": args.persistent}
for kv in args.opt:
    k, v = kv.split("=")
    opts[k] = int(v)
with g.GravityCuda(n, **opts) as h:
    if args.shard_of > 1:
        h.set_option("debug_shard_of", args.shard_of)
    h.upload(*state)
    h.step(1, 5)
    h.timer_start()
    t0 = time.perf_counter()
    h.step(1, args.steps)
    wall = (time.perf_counter() - t0) * 1e3
    ev = h.timer_stop()
    kms, kn = h.kernel_time()
    h.step(1, 1)                       # a single-step launch: its stamps describe exactly one step
    tr = h.trace_read().astype(np.int64)
info, _ = g.plan_probe(n, nranks=args.shard_of, rank=0)
t_start = tr[:, 0].min()
rel = (tr - t_start) / 1e3             # microseconds since the first CTA started


def stats(v):
    return {"min": round(float(v.min()), 2), "median": round(float(np.median(v)), 2), "max": round(float(v.max()), 2)}


pairs = float(min(n, info["n_local"])) * (n - 1)
print(json.dumps({
    "n": n, "shard_of": args.shard_of, "persistent": args.persistent, "plan": info,
    "ms_per_step_batch_events": ev / args.steps, "ms_per_step_batch_wall": wall / args.steps,
    "kernel_ms_total": kms, "kernel_launches": kn,
    "ideal_ms_at_1.27e12": pairs / 1.27e12 * 1e3,
    "single_step_launch_us": {
        "cta_start": stats(rel[:, 0]), "first_tile_ready": stats(rel[:, 1] - rel[:, 0]),
        "pairs_done_since_start": stats(rel[:, 2] - rel[:, 0]), "commit_duty": stats(rel[:, 3] - rel[:, 2]),
        "cta_end": stats(rel[:, 4]), "kernel_span": round(float(rel[:, 4].max()), 2)},
}))
