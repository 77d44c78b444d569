/*
 * include/gravity_cuda.h -- C ABI of libgravity_cuda, the B200 (sm_100a)
 * replacement for the all-pairs gravity step of sthaid/proj_entropy.
 *
 * The reference has no plugin/FFI seam for this path: the arithmetic is inline
 * in the worker thread of a static-state translation unit whose only export
 * is `void sim_gravity(void)` (util.h:42).  This header therefore DEFINES the
 * seam.  It is cut around the body of `if (local_state == STATE_RUN) {...}`
 * (sim_gravity.c:1176-1261): the k-loop (:1183-1222), barrier B2 (:1226) and
 * the NEXT_* -> state copy (:1242-1245) become one call made by worker 0;
 * `sim.sim_time += DT` (:1255) and the once-per-day PATH sampling
 * (:1232-1236, :1248-1251, :1258-1260) stay in host C.  INTEGRATION.md shows
 * the patch a maintainer of the reference would apply.
 *
 * Conventions, mirroring the reference:
 *   - plain C types only; no C++ types or exceptions cross this boundary;
 *   - every function returns 0 on success and a negative value on failure
 *     (the `ret = -1` pattern of sim_gravity.c:218,264-268) and never calls
 *     exit() (unlike FATAL, util.h:236-240); the message, at most 199
 *     characters so that it fits one error_str row (sim_gravity.c:136), is
 *     read with gravity_cuda_last_error();
 *   - all body state is fp64 SI (m, m/s, kg; sim_gravity.c:85-90); the step
 *     size is an int32 number of seconds like DT (sim_gravity.c:124) and is
 *     snapshotted per call;
 *   - G = 6.673E-11 (sim_gravity.c:29), 2-D, no softening; the evaluated body
 *     is half-drifted, the others are not (sim_gravity.c:1186-1197); self is
 *     skipped by index (:1191) and exactly coincident pairs are skipped
 *     (:1199-1201);
 *   - there is NO CPU fallback: without a usable sm_100 device every entry
 *     point fails with an error.
 *
 * A handle is used by one host thread at a time (in the reference only thread
 * 0 commits, sim_gravity.c:1229).  Host arrays are borrowed for the duration
 * of a call; device state is authoritative between upload and download.
 */
#ifndef GRAVITY_CUDA_H
#define GRAVITY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gravity_cuda gravity_cuda_t;

#define GRAVITY_CUDA_G 6.673E-11          /* sim_gravity.c:29 */

/* Which kernel family evaluates sim_gravity.c:1189-1205. */
#define GRAVITY_CUDA_KERNEL_AUTO    0     /* N <= 1024: STRICT, else TILED       */
#define GRAVITY_CUDA_KERNEL_TILED   1     /* rsqrt-based, j tiles in shared
                                             memory, work split over all SMs;
                                             matches the reference to ~1e-15     */
#define GRAVITY_CUDA_KERNEL_STRICT  2     /* IEEE sqrt/div, no FMA, j ascending:
                                             bit-identical to sim_gravity.c      */

#define GRAVITY_CUDA_NCCL_ID_BYTES 128
#define GRAVITY_CUDA_P2P_BLOB_BYTES 256

/* ---- life cycle --------------------------------------------------------- */

/* Single-process handle over devices 0..ngpus-1 of one box, 1 <= ngpus <= 8
 * (the GPUs of one NVSwitch box; typically 1, 2, 4 or 8).  Replaces worker
 * creation at sim_gravity.c:425-442: the i-bodies that the reference strides
 * over max_thread pthreads (:1183) are sharded in contiguous blocks over the
 * GPUs. */
int gravity_cuda_create(gravity_cuda_t **out, int64_t n, int ngpus);

/* One-process-per-GPU handle (torchrun style): this process is `rank` of
 * `nranks` and drives CUDA device `device`.  `nccl_id` is the
 * GRAVITY_CUDA_NCCL_ID_BYTES blob produced by gravity_cuda_nccl_unique_id() on
 * rank 0 and broadcast by the host; it may be NULL when nranks == 1. */
int gravity_cuda_nccl_unique_id(void *id_out, size_t id_bytes);
int gravity_cuda_create_rank(gravity_cuda_t **out, int64_t n, int rank, int nranks,
                             int device, const void *nccl_id, size_t id_bytes);

/* Peer-memory exchange (option "exchange" = 1).  Instead of an NCCL all-gather,
 * the step kernel stores each new position, as soon as it is committed, straight
 * into every peer's copy of the next position buffer over NVLink and publishes
 * its step number when the shard is done; in the next step (the same launch: a
 * batch of steps is one launch per GPU) every CTA fetches its own shard's
 * j-tiles first and waits for the flags only when it reaches other ranks' tiles.  Barrier B2 and the commit copy of the reference
 * (sim_gravity.c:1226, :1242-1245) thus cost no launch and overlap the math.
 * A single-process handle connects itself when the option is set.  Rank-mode
 * handles exchange one GRAVITY_CUDA_P2P_BLOB_BYTES blob per rank through the
 * host: export on every rank, all-gather the blobs in rank order, connect. */
int gravity_cuda_p2p_export(gravity_cuda_t *h, void *blob, size_t bytes);
int gravity_cuda_p2p_connect(gravity_cuda_t *h, const void *blobs_all_ranks, size_t bytes);

/* Replaces sim_gravity_terminate's worker join (sim_gravity.c:487-502).  Waits
 * for the handle's own work and, with the peer-memory exchange, until every peer
 * has published the handle's current step (a peer's last step kernel stores into
 * this device's buffers until then) before anything is freed.  In rank mode it is
 * therefore collective like step: call it on every rank after the same steps. */
void gravity_cuda_destroy(gravity_cuda_t *h);

/* Message of the last failure on this handle; h == NULL returns the message of
 * the last failed create on the calling thread. */
const char *gravity_cuda_last_error(const gravity_cuda_t *h);

/* Options: "kernel" (GRAVITY_CUDA_KERNEL_*), "threads" (CTA size of the tiled
 * kernel: 128|256), "bodies_per_thread" (1|2|4), "ctas_per_sm" (1..8) -- 0, the
 * default, lets the library pick the shape that fills the SMs best for N,
 * "exchange" (0: NCCL all-gather of the new positions, default; 1: peer-memory
 * stores + flags, see gravity_cuda_p2p_connect; collective in rank mode),
 * "mass_fold" (0|1, default 1: j-tiles whose 256 masses are bit-identical are
 * summed without the mass and scaled once per tile -- 12 instead of 13 FP64
 * instructions per pair; results differ from mass_fold=0 by rounding only),
 * "persistent" (0|1, default 1: all nsteps of a step call run in ONE launch of
 * the tiled kernel, the steps chained by per-i-block ready counters instead of
 * kernel boundaries; used on one GPU and with exchange=1; 0 = one launch per
 * step, which the NCCL exchange always uses),
 * "download_scope" (0: download/accel/path_read return all bodies on every rank,
 * default; 1: in rank mode only the caller's own shard [i0, i0+n_local) of the
 * output arrays is written -- no gather, 32*N/P bytes per rank),
 * "spin_timeout_ms" (default 10000; 0 = wait for ever): how long a kernel waits
 * for a peer's step flag or for an i-block of its own device before it gives up.
 * A wait that gives up sets a device flag that stops every kernel of the handle
 * from storing or publishing anything further; the state is then partly advanced
 * and the handle is FAIL-STOP: sync/step/download report the timeout and every
 * later call fails.  Raise the limit when a rank may legitimately stall longer
 * between two collective calls.
 * "trace" (0|1): record per-CTA time stamps of the tiled kernel (gravity_cuda_trace_read).
 * "debug_shard_of" (1..8, one-GPU handles only; developer aid): plan and run like rank 0
 * of that many ranks, without peers -- only that rank's bodies are advanced.
 * Shape options must be set before the first step/accel after create. */
int gravity_cuda_set_option(gravity_cuda_t *h, const char *key, int64_t value);

/* ---- state -------------------------------------------------------------- */

/* Host SoA, n doubles each: MASS, X, Y, VX, VY of object_t
 * (sim_gravity.c:85-90).  Every device copies only its own shard from the host
 * and the devices complete each other over NVLink (one all-gather), so 40*N
 * bytes cross PCIe in total whatever the number of GPUs.  In rank mode every
 * rank passes full-length arrays holding the same values; only the entries of
 * its own shard are read.  `mass` may be NULL after the first upload: the masses
 * stay as they are (they never change during a run, sim_gravity.c:1193).
 * Returns as soon as the host arrays have been consumed. */
int gravity_cuda_upload(gravity_cuda_t *h, const double *mass, const double *x,
                        const double *y, const double *vx, const double *vy);

/* Any pointer may be NULL.  Full-length arrays are returned on every rank unless
 * "download_scope" = 1. */
int gravity_cuda_download(gravity_cuda_t *h, double *x, double *y, double *vx, double *vy);

/* ---- the hot path ------------------------------------------------------- */

/* nsteps iterations of sim_gravity.c:1183-1222 + the commit at :1242-1245.
 * Returns after the device has finished.  The caller advances sim_time by
 * dt*nsteps (:1255).  To keep PATH sampling identical to :1234-1251 the host
 * passes nsteps only up to the next day boundary. */
int gravity_cuda_step(gravity_cuda_t *h, int32_t dt, int64_t nsteps);

/* Same, but only enqueues the work; gravity_cuda_sync() waits for it. */
int gravity_cuda_step_async(gravity_cuda_t *h, int32_t dt, int64_t nsteps);
int gravity_cuda_sync(gravity_cuda_t *h);

/* PATH ring on the device.  The reference's commit block appends every body's
 * new position to its PATH ring once per simulated day (sim_gravity.c:1230-1260:
 * `day = sim_time/86400` taken BEFORE `sim_time += DT`, a sample when it differs
 * from the function-static last_day).  path_enable() allocates a ring of
 * `capacity` rows of n positions per device (0 frees it) and resets the sample
 * count; step_path() runs nsteps like gravity_cuda_step() and lets the step
 * kernels themselves take the samples -- also inside the kernels that run a whole
 * batch in one launch -- so a run at DT = 1 day needs no host round trip per step.
 * The caller passes sim_time before the batch and last_day (in/out) and advances
 * sim_time by dt*nsteps itself; *path_tail receives the samples taken since
 * path_enable.  path_read() returns samples first..first+count-1 sample-major
 * (x[k*n + i]); only the last `capacity` samples are available. */
int gravity_cuda_path_enable(gravity_cuda_t *h, int64_t capacity);
int gravity_cuda_step_path(gravity_cuda_t *h, int32_t dt, int64_t nsteps, int64_t sim_time,
                           int32_t *last_day, int64_t *path_tail);
int gravity_cuda_path_read(gravity_cuda_t *h, int64_t first, int64_t count, double *x, double *y);

/* Host-only arithmetic (no GPU touched): how many PATH samples nsteps steps of size dt take
 * from sim_time with the given last_day -- the commit block's day test (sim_gravity.c:
 * 1232-1236) exactly as the kernels run it; *last_day is advanced.  -1 on a bad argument
 * (dt < 1, sim_time < 0, nsteps < 0).  In rank mode path_read is collective unless
 * "download_scope" = 1. */
int64_t gravity_cuda_path_samples_in(int32_t dt, int64_t nsteps, int64_t sim_time, int32_t *last_day);

/* AX, AY exactly as at sim_gravity.c:1207-1208 for the current state (n
 * doubles each); the state is not advanced. */
int gravity_cuda_accel(gravity_cuda_t *h, int32_t dt, double *ax, double *ay);

/* The same path for a host that keeps its state as SoA arrays: upload (rules of
 * gravity_cuda_upload; mass may be NULL after the first upload), nsteps steps and
 * the download back INTO x, y, vx, vy (rules of gravity_cuda_download, incl.
 * "download_scope") as one call: everything is enqueued back to back and waited
 * for once, which is what matters when a step lasts a few hundred microseconds.
 * Page-locked arrays are read and written in place by the device. */
int gravity_cuda_step_host(gravity_cuda_t *h, const double *mass, double *x, double *y,
                           double *vx, double *vy, int32_t dt, int64_t nsteps);

/* AoS convenience for the reference host: gathers MASS,X,Y,VX,VY from n
 * object_t-like records reached through objects[i] by byte offset (X@32, Y@40,
 * VX@48, VY@56, MASS@64 in sim_gravity.c:83-105), steps, and scatters
 * X,Y,VX,VY back, so the shim never needs the 1.5 MB struct definition. */
int gravity_cuda_step_objects(gravity_cuda_t *h, void **objects, int64_t n,
                              size_t off_x, size_t off_y, size_t off_vx, size_t off_vy,
                              size_t off_mass, int32_t dt, int64_t nsteps);

/* ---- diagnostics and measurement ---------------------------------------- */

/* kinetic = 1/2 sum m|v|^2, potential = -sum_{i<j} G m_i m_j / r_ij (r == 0
 * skipped) of the current, un-drifted state.  The reference computes neither;
 * the definitions follow the equations at sim_gravity.c:1145-1174. */
int gravity_cuda_energy(gravity_cuda_t *h, double *kinetic, double *potential);

/* CUDA-event timer on the handle's compute stream(s): start/stop bracket any
 * number of step_async calls; *ms is the maximum over the handle's devices. */
int gravity_cuda_timer_start(gravity_cuda_t *h);
int gravity_cuda_timer_stop(gravity_cuda_t *h, double *ms);

/* Kernels launched by this handle since create (all devices). */
int gravity_cuda_launch_count(const gravity_cuda_t *h, int64_t *launches);

/* Device time (ms) of the dominant kernel (the tiled pair-accumulation kernel,
 * or the strict one in STRICT mode) summed over the launches since the last
 * timer_start, measured with CUDA events around each launch, and how many
 * launches that was.  Valid after timer_stop. */
int gravity_cuda_kernel_time(gravity_cuda_t *h, double *ms_total, int64_t *launches);

/* The individual launch durations behind gravity_cuda_kernel_time (ms, in launch
 * order, at most `cap` of them); *count receives how many were recorded. */
int gravity_cuda_kernel_times(gravity_cuda_t *h, double *ms_out, int64_t cap, int64_t *count);

/* Developer aid (option "trace" = 1): per-CTA globaltimer stamps (ns) of the last
 * tiled launch on the handle's first device, GRAVITY_CUDA_TRACE_EVENTS per CTA:
 * [0] CTA start, [1] first j-tile in shared memory, [2] tiles of the first step
 * done, [3] commit duty of the first step done, [4] CTA end. */
#define GRAVITY_CUDA_TRACE_EVENTS 8
int gravity_cuda_trace_read(gravity_cuda_t *h, uint64_t *stamps, int64_t cap_ctas, int64_t *n_cta);

/* Dependent-free DFMA microbenchmark on `device`: measured FP64 FMA-pipe
 * throughput in TFLOP/s (2 flops per lane-FMA) and the SM clock it ran at
 * (MHz, from the device's clock64 against the event time). */
int gravity_cuda_fp64_peak(int device, double *tflops, double *sm_mhz);

/* Host-only view of the tiled kernel's launch plan for rank `rank` of `nranks`
 * on a device with `num_sms` SMs (threads/bpt/ctas_per_sm = 0: automatic).  No
 * GPU is touched, so the planner can be checked anywhere.  info[10] receives
 * {threads, bodies_per_thread, ctas_per_sm, n_cta, n_segments, n_iblock, n_tile,
 *  local_tile_begin, local_tile_end, n_local}; seg_out (may be NULL) receives up to
 * seg_cap rows of {cta, iblock, tile_begin, tile_end, partial_row} in the order the
 * CTAs execute them. */
int gravity_cuda_plan_probe(int64_t n, int nranks, int rank, int num_sms, int threads, int bpt, int ctas_per_sm,
                            int32_t *info, int32_t *seg_out, int64_t seg_cap);

/* Number of usable sm_100 devices, or a negative value. */
int gravity_cuda_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
