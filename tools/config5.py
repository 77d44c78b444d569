#!/usr/bin/env python3
"""BASELINE.json configs[4] / SURVEY.md section 8d "Config 5": synthetic Plummer
N = 1,048,576 on all GPUs of the box, 100 steps, FP64 roofline report, parity on a
1,024-row subsample of i at step 0 and after step 100.  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
      --master-port 29540 tools/config5.py [--bodies N] [--steps K]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from proj_entropy_b200.dist import Ranks

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--rows", type=int, default=1024)
ap.add_argument("--mass-fold", type=int, default=0)
ap.add_argument("--exchange", type=int, default=0)
args = ap.parse_args()

ranks = Ranks("nccl")
n, dt = args.bodies, 1
state = g.plummer(n, seed=1)
nccl_id = ranks.broadcast_bytes(g.nccl_unique_id() if ranks.rank == 0 else None, 128) if ranks.world > 1 else None
h = g.GravityCuda(n, rank=ranks.rank, nranks=ranks.world, device=ranks.local_rank, nccl_id=nccl_id,
                  kernel=g.KERNEL_TILED, mass_fold=args.mass_fold)
if ranks.world > 1 and args.exchange == 1:
    h.p2p_connect(ranks.all_gather_bytes(h.p2p_export()))
    h.set_option("exchange", 1)
h.upload(*state)
rows = np.linspace(0, n - 1, args.rows).astype(np.int64)


def parity(label, st):
    """accelerations of the current device state vs the oracle on `rows`"""
    ax, ay = h.accel(dt)
    if ranks.rank != 0:
        return None
    from oracle import oracle_py as oracle
    nthreads = max(1, min((os.cpu_count() or 2) - 1, 32))
    rax, ray = oracle.accel(*st, dt, rows=rows, nthreads=nthreads)
    tax, tay = oracle.accel_ld(*st, dt, rows=rows, nthreads=nthreads)
    mag = np.hypot(rax, ray)
    err = np.hypot(ax[rows] - rax, ay[rows] - ray) / mag
    t = np.hypot(tax, tay)
    return {"at": label, "rows": int(len(rows)), "max_rel_err_vs_reference": float(err.max()),
            "median_rel_err_vs_reference": float(np.median(err)),
            "bodies_above_1e-12": int((err > 1e-12).sum()),
            "median_rel_err_vs_long_double": float(np.median(np.hypot(ax[rows] - tax, ay[rows] - tay) / t)),
            "reference_median_rel_err_vs_long_double": float(np.median(np.hypot(rax - tax, ray - tay) / t))}


p0 = parity("step 0", state)
ke0, pe0 = h.energy()
tf, mhz = g.fp64_peak(ranks.local_rank)
h.step(dt, 2)                                   # warm-up (plans, clocks); part of the trajectory
ranks.barrier()
h.timer_start()
t0 = time.perf_counter()
h.step_async(dt, args.steps - 2)
ms = ranks.max(h.timer_stop())
wall = ranks.max(time.perf_counter() - t0)
kms, kl = h.kernel_time()
x, y, vx, vy = h.download()
p1 = parity("step %d" % args.steps, (state[0], x, y, vx, vy))
ke1, pe1 = h.energy()
if ranks.rank == 0:
    steps = args.steps - 2
    pairs = float(n) * (n - 1) * steps / (ms * 1e-3)
    print(json.dumps({
        "config": "Plummer N=%d, %d GPUs, %d steps (%d timed after 2 warm-up), DT=1" % (n, ranks.world, args.steps, steps),
        "pairs_per_s": pairs, "ms_per_step": ms / steps, "wall_s_timed": wall,
        "tflops_20_per_pair": 20 * pairs / 1e12,
        "frac_of_nominal_fp64_peak": 20 * pairs / 1e12 / (37.2 * ranks.world),
        "frac_of_measured_fp64_peak": 20 * pairs / 1e12 / (tf * ranks.world),
        "measured_fp64_peak_tflops_per_gpu": tf, "sm_mhz_during_peak_probe": mhz,
        "hot_kernel_share_of_step_rank0": kms / ms if ms else None, "hot_kernel_launches_rank0": kl,
        "parity": [p0, p1],
        "energy": {"e0": ke0 + pe0, "e1": ke1 + pe1, "relative_drift": (ke1 + pe1 - ke0 - pe0) / abs(ke0 + pe0)},
        "options": {"mass_fold": args.mass_fold, "exchange": args.exchange},
    }), flush=True)
h.close()
ranks.close()
