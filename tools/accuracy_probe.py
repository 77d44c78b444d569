#!/usr/bin/env python3
"""Developer probe: where does the tiled kernel differ most from the oracle, and
is the difference explained by the oracle's own rounding (vs long double)?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from oracle import oracle_py as oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m, x, y, vx, vy = g.plummer(n, seed=1)
with g.GravityCuda(n, kernel=g.KERNEL_TILED) as h:
    h.upload(m, x, y, vx, vy)
    ax, ay = h.accel(1)
rax, ray = oracle.accel(m, x, y, vx, vy, 1, nthreads=16)
err = np.hypot(ax - rax, ay - ray) / np.hypot(rax, ray)
order = np.argsort(err)[::-1]
worst = order[:12].astype(np.int64)
tax, tay = oracle.accel_ld(m, x, y, vx, vy, 1, rows=worst, nthreads=12)
print(json.dumps({"n": n, "max": err.max(), "median": float(np.median(err)), "p99.9": float(np.quantile(err, 0.999)),
                  "count_gt_1e-12": int((err > 1e-12).sum()), "count_gt_1e-13": int((err > 1e-13).sum())}))
amed = np.median(np.hypot(rax, ray))
for k, i in enumerate(worst):
    t = np.hypot(tax[k], tay[k])
    print(json.dumps({"i": int(i), "gpu_vs_ref": err[i], "gpu_vs_ld": float(np.hypot(ax[i] - tax[k], ay[i] - tay[k]) / t),
                      "ref_vs_ld": float(np.hypot(rax[i] - tax[k], ray[i] - tay[k]) / t), "|a|/median|a|": float(t / amed)}))
rows = np.arange(0, n, max(1, n // 4096), dtype=np.int64)
tax, tay = oracle.accel_ld(m, x, y, vx, vy, 1, rows=rows, nthreads=16)
t = np.hypot(tax, tay)
eg = np.hypot(ax[rows] - tax, ay[rows] - tay) / t
er = np.hypot(rax[rows] - tax, ray[rows] - tay) / t
print(json.dumps({"sample_rows": len(rows), "gpu_vs_ld_median": float(np.median(eg)), "gpu_vs_ld_max": float(eg.max()),
                  "ref_vs_ld_median": float(np.median(er)), "ref_vs_ld_max": float(er.max())}))
