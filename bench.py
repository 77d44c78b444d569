#!/usr/bin/env python3
"""bench.py -- fp64 pair-interactions/s of the all-pairs gravity step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--bodies B] [--impl ours|reference]

Metric (BASELINE.json): fp64 pair-interactions/s, one pair-interaction being one
(i, j != i) evaluation of sim_gravity.c:1193-1204; a step over N bodies is
N*(N-1) of them.  Workload: the synthetic planar Plummer model of SURVEY.md
section 8d at N = 1,048,576 (the size the metric is quoted on; 56 MB of state,
fits one GPU), DT = 1 s, seed 1.  N is fixed as GPUs are added: strong scaling,
i-bodies sharded over the ranks, positions all-gathered with NCCL every step.

One JSON line on stdout (rank 0):
  value     whole-job pairs/s with the state resident in HBM, K steps timed with
            CUDA events on the library's compute stream, max over ranks
  e2e       the same through the C ABI with HOST buffers: one gravity_cuda_step_host
            call per step (upload + step + download), wall clock; every rank moves
            only its own shard over PCIe (40*N bytes in, 32*N bytes out per step in total)
  parity    after the timed steps, outside the timed region: accelerations of the
            state the run ended in, on a row sample, against the oracle (CPU
            restatement of the reference loop) and against a long-double sum, and
            whether every rank holds bit-identical results
  roofline  the step kernel (pairs + reduction + kick/drift in one launch) against the FP64 FMA pipe (20 flops per pair,
            north_star's convention); the peak is measured in this run by a DFMA
            microbenchmark because MEASURED_PEAKS.json has no FP64 entry
  cpu_baseline  the reference's CPU loop (oracle port, bit-identical to the
            unmodified reference TU) on this box's host cores, row-subsampled

`--impl reference` times that CPU loop alone, on the same workload and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 pair-interactions/s"
UNIT = "pairs/s"
FLOPS_PER_PAIR = 20            # BASELINE.json north_star convention
FP64_NOMINAL_TFLOPS = 37.2     # 148 SM x 64 lanes x 2 x 1.965 GHz
# dram__bytes_read.sum + dram__bytes_write.sum of a one-step launch of k_tiled_steps<256,4,1,false> at
# the default workload (N = 1,048,576 on one GPU), from the `ncu --set full` capture summarised in
# profiles/r2_ncu_full_k_tiled_steps_1m_raw.csv.  Algorithmic bytes are 72*N = 75.5 MB (positions,
# masses, velocities read once: 40*N; new positions and velocities written: 32*N) plus the partial-sum rows
# (written and read once each, hi + lo: 64 bytes per body and segment, ~1,170 segments of 1,024 bodies at this size).
# Other configurations: not captured.
NCU_DRAM_BYTES_PER_STEP = {(1 << 20, 1): 72227328 + 58004736}
DT = 1
SEED = 1


def workload_config(n):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "planar Plummer model N=%d, DT=%d s, seed %d (BASELINE.json configs[4] size at N=1048576)" % (n, DT, SEED),
            "bodies": n, "dt_s": DT, "seed": SEED,
            "l2": "compute-bound: 24*N bytes of j-state stream from L2/HBM per i-block, no flush needed; "
                  "inputs change every step"}


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line.  Native libraries print there too
    (NCCL announces its version on stdout when a communicator is created), so
    file descriptor 1 is pointed at stderr for the rest of the run and the real
    stdout is kept aside for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(record):
    data = (json.dumps(record) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bodies", type=int, default=1 << 20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(--steps, 5), or 50 up to 262,144 bodies")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--bodies-per-thread", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--exchange", type=int, default=1,
                    help="multi-GPU position exchange: 0 = NCCL all-gather, 1 = peer-memory stores from the step kernel")
    ap.add_argument("--mass-fold", type=int, default=0,
                    help="headline kernel path: 0 = general masses (13 FP64 instr/pair, default), 1 = let equal-mass "
                         "j-tiles take the 12-instruction path")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational equal-mass-path measurement")
    ap.add_argument("--persistent", type=int, default=1,
                    help="1 = all timed steps in one launch of the step kernel (default), 0 = one launch per step")
    ap.add_argument("--parity-rows", type=int, default=1024, help="rows of the in-run parity sample (0: skip)")
    ap.add_argument("--spin-timeout-ms", type=int, default=-1,
                    help="library option spin_timeout_ms (-1: library default, 10 s; 0: wait for ever -- use under "
                         "ncu --set full, whose instrumented passes slow the kernel far beyond any sane limit)")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# CPU leg: the reference's loop on the host cores
# ---------------------------------------------------------------------------

def host_threads():
    """The reference's own policy: min(nproc - 1, 16) workers (sim_gravity.c:425-432)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, min(n - 1, 16)), n


def cpu_sample(state, rows, nthreads):
    """Time `rows` i-bodies against all N j-bodies with the oracle port; returns
    (pairs/s, seconds).  Per-pair cost does not depend on which rows are taken."""
    import numpy as np
    from oracle import oracle_py as oracle
    m, x, y, vx, vy = state
    n = len(m)
    sel = np.linspace(0, n - 1, rows).astype(np.int64)
    t0 = time.perf_counter()
    oracle.accel(m, x, y, vx, vy, DT, rows=sel, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return rows * (n - 1) / dt, dt


def all_cores_sample(state, rows, ncpu):
    """For information: the same loop on every host cpu."""
    rate, secs = cpu_sample(state, max(ncpu, rows // 2), ncpu)
    return {"value": rate, "cores": ncpu, "seconds": secs}


def reference_tu_sample(steps=200000):
    """For information: the UNMODIFIED reference translation unit (oracle/_ref,
    built from /root/reference by oracle/Makefile) at a scale it can represent --
    the authored 15-body scenario, 15 worker threads + the metering participant,
    two spin barriers per step (sim_gravity.c:1140, :1226).  Not the bench
    workload (MAX_OBJECT = 200): it shows the regime the reference really runs in."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gravity")
    scen = os.path.join(ROOT, "tests", "golden", "scenarios", "random15")
    if not (os.path.exists(exe) and os.path.exists(scen)):
        return None
    try:
        t0 = time.perf_counter()
        subprocess.run([exe, scen, "0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=60, check=True)
        t1 = time.perf_counter()
        subprocess.run([exe, scen, "0", str(steps)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300,
                       check=True)
        t2 = time.perf_counter()
    except Exception:
        return None
    secs = max(1e-6, (t2 - t1) - (t1 - t0) - 0.05)      # minus start-up and the extra 50 ms STOP round of the second dump
    return {"bodies": 15, "steps": steps, "seconds": secs, "value": 15 * 14 * steps / secs, "unit": UNIT,
            "threads": 16, "what": "unmodified sim_gravity.c through oracle/_ref/ref_gravity"}


def auto_rows(n, nthreads, target_cpu_seconds=20.0):
    # ~2e8 pairs/s/thread on current Xeons
    rows = int(target_cpu_seconds * 2.0e8 / max(1, n - 1))
    rows = max(nthreads, min(n, rows))
    return rows


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the
    host cores.  The unmodified translation unit cannot hold this workload
    (MAX_OBJECT = 200, 1.5 MB per object_t; sim_gravity.c:35,83-105), so the arm
    times the oracle port, which tests/ prove bit-identical to it."""
    if rank != 0:
        return
    claim_stdout()
    import proj_entropy_b200 as g
    n = args.bodies
    nthreads, ncpu = host_threads()
    state = g.plummer(n, seed=SEED)
    rows = args.cpu_rows or auto_rows(n, nthreads, 12.0)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_sample(state, max(nthreads, rows // 8), nthreads)
    t_total, pairs = 0.0, 0.0
    for _ in range(args.steps):
        rate, secs = cpu_sample(state, rows, nthreads)
        t_total += secs
        pairs += rows * (n - 1)
    value = pairs / t_total
    all_cores = all_cores_sample(state, rows, ncpu) if ncpu != nthreads else None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # what was really timed: one bounded sample of the workload per step
        "ms_per_step": 1e3 * t_total / args.steps,
        "full_step_ms_extrapolated": 1e3 * n * (n - 1) / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "each timed step = %d of the %d i-rows x all %d j-bodies (the per-pair cost does not "
                                   "depend on the row); %d steps; host has %d cpus; thread count is the reference's "
                                   "policy min(nproc-1,16); oracle port, bit-identical to the unmodified reference TU"
                                   % (rows, n, n, args.steps, ncpu),
                         "all_cores": all_cores, "reference_tu_small": reference_tu_sample()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.samples, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [s.split(", ") for t, s in self.samples if t0 <= t <= t1] or [s.split(", ") for _, s in self.samples[-3:]]
        sm, reasons, smmax, power = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); smmax = float(r[1]); power.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smmax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def parity_record(h, g, ranks, state_mass, rows_wanted, rank, world):
    """Outside the timed region: accelerations of the state the run ended in (sim_gravity.c:
    1189-1208), on a row sample, against the oracle and against a long-double sum (rank 0),
    plus whether every rank holds bit-identical positions, velocities and accelerations."""
    import hashlib
    import numpy as np
    n = len(state_mass)
    h.set_option("download_scope", 0)
    x, y, vx, vy = h.download()                   # full arrays on every rank
    ax, ay = h.accel(DT)                          # gathered from every rank
    digest = hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in (x, y, vx, vy, ax, ay))).digest()
    all_digests = ranks.all_gather_bytes(digest)
    identical = all(all_digests[k * 32:(k + 1) * 32] == digest for k in range(world))
    if rank != 0:
        return None
    from oracle import oracle_py as oracle
    nthreads, _ = host_threads()
    rows = np.unique(np.linspace(0, n - 1, max(512, rows_wanted)).astype(np.int64))
    rax, ray = oracle.accel(state_mass, x, y, vx, vy, DT, rows=rows, nthreads=nthreads)
    tax, tay = oracle.accel_ld(state_mass, x, y, vx, vy, DT, rows=rows, nthreads=nthreads)
    mag, tmag = np.hypot(rax, ray), np.hypot(tax, tay)
    err = np.hypot(ax[rows] - rax, ay[rows] - ray) / mag
    floor = 0.1 * float(np.median(mag))
    err_floor = np.hypot(ax[rows] - rax, ay[rows] - ray) / np.maximum(mag, floor)
    gpu_ld = np.hypot(ax[rows] - tax, ay[rows] - tay) / tmag
    ref_ld = np.hypot(rax - tax, ray - tay) / tmag
    return {"rows": int(len(rows)), "state": "after the timed and end-to-end steps", "max_rel_vs_ref": float(err.max()),
            "median_rel_vs_ref": float(np.median(err)), "n_above_1e-12": int((err > 1e-12).sum()),
            "max_rel_vs_ref_with_floor": float(err_floor.max()), "floor_used": bool((err > 1e-12).any()),
            "max_rel_vs_long_double": float(gpu_ld.max()), "ref_max_rel_vs_long_double": float(ref_ld.max()),
            "median_rel_vs_long_double": float(np.median(gpu_ld)),
            "ref_median_rel_vs_long_double": float(np.median(ref_ld)),
            "rows_where_gpu_is_farther_from_long_double_than_ref": int((gpu_ld > ref_ld).sum()),
            "ranks_bit_identical": bool(identical), "ranks_compared": world,
            "gate": "per body |a - a_ref|_2 / |a_ref|_2 <= 1e-12 (BASELINE.json); tests/conftest.py:assert_accel_parity "
                    "additionally tolerates, for at most 1 body in 10^4 whose net force nearly cancels, the same bound "
                    "taken against max(|a_ref|, 0.1 median|a_ref|) if that body is no farther from a long-double sum "
                    "than 4x the reference's own distance; floor_used says whether this sample needed it",
            "pass": bool(err_floor.max() <= 1e-12 and (err > 1e-12).sum() <= max(1, len(rows) // 10000) and identical)}


def run_ours(args, rank, world, local_rank, ranks=None):
    """ranks: an open proj_entropy_b200.dist.Ranks to reuse (tools/matrix.py runs several cells in
    one process); by default one is opened here and closed at the end."""
    claim_stdout()
    import numpy as np
    import torch
    import proj_entropy_b200 as g

    if g.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 GPU visible; libgravity_cuda has no CPU fallback")
    from proj_entropy_b200.dist import Ranks, shard_bounds
    own_ranks = ranks is None
    if own_ranks:
        ranks = Ranks("nccl")
    n = args.bodies
    state = g.plummer(n, seed=SEED)        # identical bytes on every rank
    barrier, max_over_ranks, sum_over_ranks = ranks.barrier, ranks.max, ranks.sum

    # the library's own NCCL communicator: id made on rank 0, broadcast by torch
    nccl_id = None
    if world > 1:
        nccl_id = ranks.broadcast_bytes(g.nccl_unique_id() if rank == 0 else None, 128)

    opts = {"kernel": g.KERNEL_TILED, "mass_fold": args.mass_fold, "persistent": args.persistent}
    if args.spin_timeout_ms >= 0:
        opts["spin_timeout_ms"] = args.spin_timeout_ms
    if args.threads:
        opts["threads"] = args.threads
    if args.bodies_per_thread:
        opts["bodies_per_thread"] = args.bodies_per_thread
    if args.ctas_per_sm:
        opts["ctas_per_sm"] = args.ctas_per_sm
    h = g.GravityCuda(n, rank=rank, nranks=world, device=local_rank, nccl_id=nccl_id, **opts)
    exchange = 0
    if world > 1 and args.exchange == 1:
        # peer-memory exchange; if the peers cannot be mapped every rank stays on
        # the NCCL all-gather (still GPU-to-GPU over NVLink; recorded in the line)
        try:
            h.p2p_connect(ranks.all_gather_bytes(h.p2p_export()))
            ok = 1.0
        except g.GravityError as e:
            sys.stderr.write("rank %d: peer-memory exchange unavailable (%s)\n" % (rank, e))
            ok = 0.0
        if ranks.sum(ok) == world:
            h.set_option("exchange", 1)
            exchange = 1

    fp64_tf, fp64_mhz = g.fp64_peak(local_rank) if rank == 0 else (None, None)

    # ---- device-resident throughput: K steps, one call ---------------------------
    h.upload(*state)
    h.step_async(DT, args.warmup)
    h.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    torch.cuda.synchronize()
    launches0 = h.launch_count()
    t0 = time.perf_counter()
    h.timer_start()
    h.step_async(DT, args.steps)
    ms = h.timer_stop()
    torch.cuda.synchronize()
    barrier()
    t1 = time.perf_counter()
    launches = h.launch_count() - launches0
    kernel_ms, kernel_launches = h.kernel_time()
    ms = max_over_ranks(ms)
    total_launches = int(sum_over_ranks(float(launches)))
    clocks = sampler.summary(t0, t1) if rank == 0 else None
    pairs_per_step = float(n) * float(n - 1)
    value = pairs_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------
    # every step: host -> device of that step's inputs, one step, device -> host of its
    # result.  Each rank moves only its own shard over PCIe (upload reads nothing else,
    # download_scope = 1 writes nothing else); the devices complete each other over NVLink.
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 50 if n <= (1 << 18) else 5))    # bounded: minutes at most
    lo, hi = shard_bounds(n, world, rank)
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in state]
    ptrs = [t.data_ptr() for t in pinned]
    if world > 1:
        h.set_option("download_scope", 1)
    h.step_host_ptrs(*ptrs, DT, 1)                           # warm the path once
    barrier()
    torch.cuda.synchronize()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        # one C-ABI call per step: host arrays in, one step, host arrays out (updated in place, so
        # the next step's input is this step's output)
        h.step_host_ptrs(*ptrs, DT, 1)
    torch.cuda.synchronize()
    barrier()
    te = max_over_ranks(time.perf_counter() - te0)
    e2e_value = pairs_per_step * e2e_steps / te
    h2d_total = int(sum_over_ranks(float(5 * 8 * (hi - lo))))
    d2h_total = int(sum_over_ranks(float(4 * 8 * (hi - lo))))

    # ---- parity of the state this run ended in (not timed) ----------------------------
    parity = parity_record(h, g, ranks, state[0], args.parity_rows, rank, world) if args.parity_rows > 0 else None

    # ---- informational: the equal-mass fast path on the same workload ----------
    extra = None
    if not args.no_extra and not args.mass_fold:
        h.set_option("mass_fold", 1)
        h.upload(*state)
        h.step_async(DT, min(2, args.warmup))
        h.sync()
        barrier()
        h.timer_start()
        fold_steps = max(1, min(args.steps, 5))
        h.step_async(DT, fold_steps)
        ms_fold = max_over_ranks(h.timer_stop())
        barrier()
        extra = {"value": pairs_per_step * fold_steps / (ms_fold * 1e-3), "unit": UNIT,
                 "ms_per_step": ms_fold / fold_steps, "steps": fold_steps,
                 "note": "mass_fold=1: the Plummer workload has equal masses, so every j-tile takes the "
                         "12-instruction path (mass applied once per tile); NOT the headline value"}

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads, ncpu = host_threads()
        rows = args.cpu_rows or auto_rows(n, nthreads)
        rate, secs = cpu_sample(state, rows, nthreads)
        cpu = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": "%d i-rows x all %d j-bodies (%.1f s wall on %d of %d host cpus); oracle port, bit-identical "
                         "to the unmodified reference TU" % (rows, n, secs, nthreads, ncpu),
               "all_cores": all_cores_sample(state, rows, ncpu) if ncpu != nthreads else None}

    if rank == 0:
        peaks = measured_peaks()
        n_hot = max(1, kernel_launches)
        avg_kernel_ms = kernel_ms / n_hot
        steps_per_launch = args.steps / n_hot
        per_launch_pairs = pairs_per_step / world * steps_per_launch     # every rank runs 1/world of the pairs of a step
        achieved = FLOPS_PER_PAIR * per_launch_pairs / (avg_kernel_ms * 1e-3) / 1e12
        dram = NCU_DRAM_BYTES_PER_STEP.get((n, world))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n),
            "variant": {"parallelism": "one GPU, no exchange" if world == 1 else
                                       "i-sharded x%d, %s" % (world, "peer-memory stores of the new positions from the step "
                                       "kernel + step flags" if exchange == 1 else
                                       "NCCL all-gather of positions per step"),
                        "kernel": "tiled rsqrt, general masses, 13 FP64-pipe instr/pair" if not args.mass_fold else
                                  "tiled rsqrt, equal-mass tiles folded, 12 FP64-pipe instr/pair",
                        "mass_fold": args.mass_fold, "exchange": exchange, "persistent": args.persistent,
                        "steps_per_launch": steps_per_launch},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": d2h_total, "steps": e2e_steps,
                    "note": "sum over ranks: every rank copies only its own shard (40 and 32 bytes per body)"},
            "gpu_launches": total_launches,
            "parity": parity,
            "roofline": {"bound": "fp64_fma_pipe", "achieved": achieved, "peak": fp64_tf, "unit": "TFLOP/s",
                         "frac": achieved / fp64_tf if fp64_tf else None,
                         "traffic": dram * steps_per_launch if dram else None,
                         "traffic_note": "ncu dram bytes of a one-step launch x steps per launch",
                         "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry: "
                                        "keys %s)" % sorted(peaks.keys())[:4],
                         "peak_sm_mhz": fp64_mhz, "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
                         "flops_per_pair": FLOPS_PER_PAIR, "reference_fp64_ops_per_pair": 13,
                         "sass_fp64_instr_per_pair": 12 if args.mass_fold else 13, "sass_mufu_per_pair": 1,
                         "kernel": "k_tiled_steps",
                         "kernel_ms_avg": avg_kernel_ms, "kernel_launches_timed": kernel_launches,
                         "kernel_share_of_step": kernel_ms / ms if ms else None,
                         "job_frac_of_nominal": FLOPS_PER_PAIR * value / 1e12 / (FP64_NOMINAL_TFLOPS * world)},
        }
        if extra:
            line["equal_mass_path"] = extra
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    sampler.stop()
    h.close()
    if own_ranks:
        ranks.close()
    else:
        ranks.barrier()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
