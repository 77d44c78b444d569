#!/usr/bin/env python3
"""The cells of the throughput matrix for ONE GPU count in one process per rank: calls
bench.py's own run_ours() once per (bodies, exchange) cell, so that interpreter start-up,
the torch import and the Plummer generator's libm are paid once instead of once per cell.
Every cell's line is exactly the line `bench.py --bodies B --exchange X ...` prints.

usage (NGPU > 1 under torchrun, like bench.py):
    python tools/matrix.py --gpus NGPU --out OUTDIR [--prefix r2] [--sizes 65536,262144,1048576] [--exchanges 1,0]
writes OUTDIR/<prefix>_scale_g<NGPU>_n<B>_x<X>.json and prints a one-line summary per cell.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default="gpurun_out")
ap.add_argument("--prefix", default="r2")
ap.add_argument("--sizes", default="65536,262144,1048576")
ap.add_argument("--exchanges", default="1,0")
cli = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
os.makedirs(cli.out, exist_ok=True)
exchanges = [int(x) for x in cli.exchanges.split(",")] if world > 1 else [1]
from proj_entropy_b200.dist import Ranks
ranks = Ranks("nccl")                                  # one rendezvous for all cells

for bodies in [int(b) for b in cli.sizes.split(",")]:
    steps = {65536: 200, 262144: 40}.get(bodies, 10 if world >= 4 else 5)
    for x in exchanges:
        sys.argv = ["bench.py", "--gpus", str(world), "--steps", str(steps), "--warmup", "3", "--bodies", str(bodies),
                    "--exchange", str(x), "--no-cpu-baseline"]
        args = bench.parse_args()
        lines = []
        bench.emit = lines.append                      # rank 0 emits one record
        bench.run_ours(args, rank, world, local_rank, ranks=ranks)
        if rank == 0 and lines:
            d = lines[0]
            path = os.path.join(cli.out, "%s_scale_g%d_n%d_x%d.json" % (cli.prefix, world, bodies, x))
            json.dump(d, open(path, "w"))
            p = d.get("parity") or {}
            sys.stderr.write("gpus=%d n=%d x=%s pairs/s=%.4g ms/step=%.4f frac_nominal=%.3f e2e=%.4g (%.0f%% of resident) "
                             "parity max=%.2g above=%s identical=%s\n" % (
                                 d["n_gpus"], bodies, d["variant"]["exchange"], d["value"], d["ms_per_step"],
                                 d["roofline"]["job_frac_of_nominal"], d["e2e"]["value"], 100 * d["e2e"]["value"] / d["value"],
                                 p.get("max_rel_vs_ref", float("nan")), p.get("n_above_1e-12"), p.get("ranks_bit_identical")))
ranks.close()
