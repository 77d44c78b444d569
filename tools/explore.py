#!/usr/bin/env python3
"""Developer sweep (not the bench): time the tiled kernel shapes at a few N and
print pair-interactions/s against the in-run DFMA peak."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g


def time_steps(h, dt, steps):
    h.step(dt, 2)
    h.timer_start()
    h.step_async(dt, steps)
    ms = h.timer_stop()
    kms, kl = h.kernel_time()
    return ms / steps, (kms / kl if kl else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="65536,262144")
    ap.add_argument("--shapes", default="128x1x4,128x2x4,128x2x3,128x4x2,128x4x3,256x1x2,256x2x2,256x2x1,256x4x1,256x4x2")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    tf, mhz = g.fp64_peak(0)
    print(json.dumps({"fp64_peak_tflops": tf, "sm_mhz": mhz}), flush=True)
    for n in [int(s) for s in args.sizes.split(",")]:
        m, x, y, vx, vy = g.plummer(n, seed=1)
        for shape in args.shapes.split(","):
            opts = {}
            if shape != "auto":
                t, b, c = [int(v) for v in shape.split("x")]
                opts = dict(threads=t, bodies_per_thread=b, ctas_per_sm=c)
            with g.GravityCuda(n, kernel=g.KERNEL_TILED, mass_fold=0, **opts) as h:
                h.upload(m, x, y, vx, vy)
                steps = max(1, min(args.steps, int(3e11 / (n * n)) + 1))
                ms, kms = time_steps(h, 1, steps)
                pairs = n * (n - 1) / (ms * 1e-3)
                print(json.dumps({"n": n, "shape": shape, "ms_per_step": round(ms, 4), "kernel_ms": round(kms, 4),
                                  "pairs_per_s": pairs, "tflops20": 20 * pairs / 1e12,
                                  "frac_of_measured_peak": 20 * pairs / 1e12 / tf}), flush=True)


if __name__ == "__main__":
    main()
