#!/usr/bin/env python3
"""Developer probe: how far are the tiled kernel's accelerations from the oracle, and how far
are both from a long-double evaluation of the same sums?  Every body is checked against the
long-double yardstick (x87, threaded), so the "worst body" is not a selection effect.

usage: python tools/accuracy_probe.py [N] [--mass-fold 0|1]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import proj_entropy_b200 as g
from oracle import oracle_py as oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 65536
fold = int(sys.argv[sys.argv.index("--mass-fold") + 1]) if "--mass-fold" in sys.argv else 0
nthreads = max(1, min((os.cpu_count() or 2) - 1, 32))
m, x, y, vx, vy = g.plummer(n, seed=1)
with g.GravityCuda(n, kernel=g.KERNEL_TILED, mass_fold=fold) as h:
    h.upload(m, x, y, vx, vy)
    ax, ay = h.accel(1)
rows = np.arange(n, dtype=np.int64) if n <= 131072 else np.linspace(0, n - 1, 8192).astype(np.int64)
rax, ray = oracle.accel(m, x, y, vx, vy, 1, rows=rows, nthreads=nthreads)
tax, tay = oracle.accel_ld(m, x, y, vx, vy, 1, rows=rows, nthreads=nthreads)
t = np.hypot(tax, tay)
gpu_ref = np.hypot(ax[rows] - rax, ay[rows] - ray) / np.hypot(rax, ray)
gpu_ld = np.hypot(ax[rows] - tax, ay[rows] - tay) / t
ref_ld = np.hypot(rax - tax, ray - tay) / t
q = lambda v, p: float(np.quantile(v, p))
print(json.dumps({
    "n": n, "rows_checked": int(len(rows)), "mass_fold": fold,
    "gpu_vs_ref": {"max": float(gpu_ref.max()), "median": q(gpu_ref, .5), "p99.9": q(gpu_ref, .999),
                   "count_gt_1e-12": int((gpu_ref > 1e-12).sum()), "count_gt_1e-13": int((gpu_ref > 1e-13).sum())},
    "gpu_vs_long_double": {"max": float(gpu_ld.max()), "median": q(gpu_ld, .5), "p99.9": q(gpu_ld, .999)},
    "ref_vs_long_double": {"max": float(ref_ld.max()), "median": q(ref_ld, .5), "p99.9": q(ref_ld, .999)},
    "bodies_where_gpu_is_farther_from_long_double_than_ref": int((gpu_ld > ref_ld).sum()),
    "bodies_where_gpu_is_farther_and_gpu_vs_ld_gt_1e-14": int(((gpu_ld > ref_ld) & (gpu_ld > 1e-14)).sum()),
    "note": "the long-double sum itself carries ~1e-19 relative rounding per addition; differences below ~1e-16 are ties"}))
amed = np.median(t)
for i in np.argsort(gpu_ref)[::-1][:8]:
    print(json.dumps({"i": int(rows[i]), "gpu_vs_ref": float(gpu_ref[i]), "gpu_vs_ld": float(gpu_ld[i]),
                      "ref_vs_ld": float(ref_ld[i]), "|a|/median|a|": float(t[i] / amed)}))
