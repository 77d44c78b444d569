#!/usr/bin/env python3
"""One-GPU sweep of the tiled step over N (SURVEY.md section 7 hard part 5): for each N
one gravity_cuda_step(h, dt, K) batch is timed three ways -- host wall clock around
the call, CUDA events around the batch on the library's stream, and the sum of the
per-launch kernel durations -- so that launch gaps and per-step tails show up as the
difference between the three.  One JSON line per N.

usage: python tools/n_sweep.py [--sizes 1025,2048,...] [--label before|after] [--kernel tiled|strict]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proj_entropy_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1025,1536,2048,3072,4096,6144,8192,12288,16384,24576,32768,65536,131072,262144,1048576")
ap.add_argument("--label", default="")
ap.add_argument("--kernel", default="tiled")
ap.add_argument("--opt", action="append", default=[], help="key=value passed to gravity_cuda_set_option")
ap.add_argument("--target-ms", type=float, default=300.0)
args = ap.parse_args()

PAIR_RATE = 1.27e12          # pairs/s of the large-N kernel on one B200 (profiles/r1_summary.md)
for n in [int(s) for s in args.sizes.split(",")]:
    state = g.plummer(n, seed=1)
    opts = {"kernel": g.KERNEL_TILED if args.kernel == "tiled" else g.KERNEL_STRICT}
    for kv in args.opt:
        k, v = kv.split("=")
        opts[k] = int(v)
    with g.GravityCuda(n, **opts) as h:
        h.upload(*state)
        ideal_ms = n * (n - 1.0) / PAIR_RATE * 1e3
        steps = int(max(3, min(2000, args.target_ms / max(ideal_ms, 0.005))))
        h.step(1, min(steps, 20))                      # warm-up: plan, clocks
        h.timer_start()
        t0 = time.perf_counter()
        h.step(1, steps)
        wall_ms = (time.perf_counter() - t0) * 1e3
        ev_ms = h.timer_stop()
        k_ms, k_n = h.kernel_time()
        # one step at a time, synchronised: the cost a caller pays per call
        t0 = time.perf_counter()
        for _ in range(min(steps, 50)):
            h.step(1, 1)
        single_ms = (time.perf_counter() - t0) * 1e3 / min(steps, 50)
    print(json.dumps({"label": args.label, "n": n, "steps": steps, "kernel": args.kernel,
                      "wall_ms_per_step": wall_ms / steps, "event_ms_per_step": ev_ms / steps,
                      "kernel_ms_per_launch": (k_ms / k_n) if k_n else None, "kernel_launches": k_n,
                      "sync_step_ms": single_ms, "ideal_ms_at_large_n_rate": ideal_ms,
                      "pairs_per_s": n * (n - 1.0) * steps / (ev_ms * 1e-3),
                      "efficiency_vs_large_n": ideal_ms / (ev_ms / steps)}), flush=True)
