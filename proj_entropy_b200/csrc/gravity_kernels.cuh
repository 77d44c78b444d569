// gravity_kernels.cuh -- sm_100a device code of libgravity_cuda.
//
// Everything here restates, for the GPU, the arithmetic that is inline in the
// reference worker thread (sim_gravity.c:1183-1222).  Reference line numbers
// are cited next to the code that carries the same meaning.
//
// Data layout in HBM (per device):
//   pos[2][npad]   double2 {X,Y}, ping-pong (time t / t+1), ALL bodies replicated
//   mass[npad]     double, replicated, immutable after upload
//   vel[chunk]     double2 {VX,VY}, only this device's i-shard
//   partial[nseg][IB] double2, per work-segment partial sums of the tiled kernel
// Bodies with global index >= n are padding: mass 0, parked far away, never
// updated; they contribute exactly 0 to every sum.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gravity {

constexpr int    kTile        = 256;     // j-bodies per shared-memory tile
#ifndef GRAVITY_PAIRS_GENERAL_R4
#define GRAVITY_PAIRS_GENERAL_R4 32
#endif
#ifndef GRAVITY_PAIRS_FOLD_R4
#define GRAVITY_PAIRS_FOLD_R4 32
#endif
constexpr int    kPairsPerIteration = 16;           // pairs per unrolled iteration of the pair loop ...
constexpr int    kPairsPerIterationFoldR4 = GRAVITY_PAIRS_FOLD_R4;
constexpr int    kPairsPerIterationGeneralR4 = GRAVITY_PAIRS_GENERAL_R4;   // ... and of its general-mass, 4-bodies-per-thread form
constexpr int    kStages      = 4;       // tile ring depth
constexpr int    kLookahead   = 2;       // tiles in flight ahead of the consumer
constexpr int    kSmallMax    = 1024;    // largest N of the single-CTA kernel
constexpr double kG           = 6.673E-11;   // sim_gravity.c:29
constexpr double kPadCoord    = 1.0e150;     // parking spot of padding bodies
constexpr int    kMaxPeers    = 8;           // ranks of the peer-memory exchange (one NVSwitch box)
constexpr int    kTraceEvents = 8;           // per-CTA time stamps of the optional trace

// status word values (0 = fine).  Once it is non-zero every kernel of the handle
// stops storing and publishing: the handle is fail-stop.
constexpr unsigned int kStatusPeerTimeout  = 1u;   // a peer never published its step
constexpr unsigned int kStatusLocalTimeout = 2u;   // an i-block of this device never got ready

// The commit block's clock (sim_gravity.c:1232-1236, :1255-1260), carried into the
// kernels that run many steps per launch so that the once-per-simulated-day PATH
// sample needs no host round trip.  `day`/`rem` describe sim_time BEFORE the
// step's `sim_time += DT`, which is what the reference tests.
struct StepClock {
    long long day;        // sim_time / 86400
    long long rem;        // sim_time % 86400
    long long tail;       // samples taken so far (path_tail); ring row = tail % capacity
    long long capacity;   // rows of the ring; 0 = PATH sampling off
    double2  *ring;       // [capacity][npad] positions, one row per sample
    long long npad;
    int32_t   last_day;   // function-static in the reference
    int32_t   dt;         // > 0
};

// One commit's worth of clock: returns the ring row to sample into, or -1.
__host__ __device__ inline long long clock_tick(StepClock &c)
{
    long long row = -1;
    if (c.capacity > 0) {
        const int32_t day = (int32_t)c.day;                      // `int32_t day = sim.sim_time/86400` (:1234)
        if (day != c.last_day) row = c.tail++ % c.capacity;      // :1235, :1248-1251, :1258-1260
        c.last_day = day;                                        // :1236
        c.rem += c.dt;                                           // :1255
        if (c.rem >= 86400) {
            const long long q = c.rem / 86400;
            c.day += q;
            c.rem -= q * 86400;
        }
    }
    return row;
}

// Peer-memory exchange (replaces the NCCL all-gather when connected): the step
// kernel stores every new position into each peer's copy of pos_next over
// NVLink as soon as it is committed and, when the rank's last i-block is done,
// publishes its step number in each peer's flag array; before a kernel fetches
// the first tile of other ranks' bodies for a step it waits until every peer's
// flag has reached that step.
struct PeerExchange {
    double2            *peer_pos[2][kMaxPeers];  // both position buffers of every rank, as mapped here
    unsigned long long *peer_flags[kMaxPeers];   // flag array of every rank, as mapped here
    unsigned long long *flags;                   // this rank's own flag array [kMaxPeers]
    int32_t             rank, nranks;
    int32_t             enabled;
    int32_t             reserved;
};

// One contiguous run of j-tiles for one block of i-bodies, executed by one CTA.
struct Segment {
    int32_t iblock;       // which block of IB = THREADS*R local i-bodies
    int32_t tile_begin;   // first j-tile (units of kTile bodies)
    int32_t tile_end;     // one past the last j-tile
    int32_t slot;         // which partial[] row this segment writes
};

// Spin-wait policy shared by every kernel that waits on another CTA or GPU.
struct WaitPolicy {
    unsigned int *status;          // device word, see kStatus*: what the kernels poll
    unsigned int *host_status;     // the same verdict in mapped host memory: what the host reads, without a copy
    long long     timeout_cycles;  // 0: wait for ever
};

// Parameters of the strict row kernel (one launch per step).
struct CommitParams {
    const double2 *pos_cur;         // time-t positions (all bodies)
    double2       *pos_next;        // time-t+1 positions (all bodies; we write our shard)
    double2       *vel;             // local shard, updated in place
    const double  *mass;
    double2       *accel_out;       // local shard, used when export_accel
    double2       *ring_row;        // PATH sample row for this step, or NULL
    int64_t        n;               // real bodies, all shards
    int64_t        i_global0;
    int64_t        n_local;
    int32_t        export_accel;    // 1: write AX,AY and leave the state alone
    int32_t        cur_next;        // index of pos_next among the two buffers (peer stores)
    double         dt;
    unsigned int  *commit_counter;  // last-CTA detection
    unsigned long long wait_step;   // peers must have published this step before pos_cur is read
    unsigned long long publish_step;
    WaitPolicy     wp;
    PeerExchange   px;
};

// Parameters of the tiled step kernel.
struct TiledParams {
    double2       *pos[2];          // ping-pong positions of ALL bodies; pos[cur] is time t at step 0
    const double  *mass;
    double2       *vel;             // local shard, updated in place
    double2       *accel_out;       // local shard, used when export_accel
    double2       *partial_hi;      // [nseg][IB] per-segment partial sums ...
    double2       *partial_lo;      // ... and their compensation terms
    const Segment *seg;
    const int32_t *cta_seg_begin;   // [gridDim.x + 1]
    const int32_t *iblock_seg_begin;    // [n_iblock + 1] partial rows of each i-block
    const double  *tile_mass;       // [n_tile] common mass of a j-tile whose 256 masses are
                                    // bit-identical, NaN otherwise; NULL = never fold
    // dataflow counters, all monotonic (never reset while the plan lives)
    unsigned long long *arrivals;       // [n_iblock] partial rows written, all evaluations
    unsigned long long *slices_done;    // [n_iblock] slices committed, all evaluations
    unsigned long long *iblock_ready;   // [n_iblock] absolute step number this i-block has reached
    unsigned long long *xfer_done;      // [1] slices whose stores to the peers are visible system-wide (committed steps only)
    unsigned long long  xfer_base;      // value of *xfer_done before this launch
    int32_t             n_seg_total;    // segments (= slices) of one step on this device
    int32_t             reserved2;
    unsigned long long *trace;          // [gridDim.x][kTraceEvents] globaltimer stamps, or NULL
    unsigned long long  epoch;          // evaluations done with this plan before this launch
    unsigned long long  step_base;      // absolute step number of the state at launch
    long long           nsteps;         // steps in this launch (1 when export_accel)
    int64_t        n;               // real bodies, all shards
    int64_t        i_global0;       // global index of local body 0
    int64_t        n_local;         // real bodies in the local shard
    int32_t        n_iblock;
    int32_t        cur;
    int32_t        export_accel;
    int32_t        local_tile_begin, local_tile_end;   // j-tiles of this rank's own shard
    int32_t        reserved;
    double         dt;              // (double)DT
    StepClock      clock;
    WaitPolicy     wp;
    PeerExchange   px;
};

// ---------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------
// Waiting on other CTAs and other GPUs
// ---------------------------------------------------------------------------

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin (one thread) until *p >= target.  Returns false -- after recording `code`
// in the status word -- when the wait timed out, and at once when another waiter
// of this handle already gave up, so that one missing producer never hangs the
// GPU and never lets a kernel commit stale data: callers stop storing on false.
template <bool SYS>
__device__ __forceinline__ bool spin_until_reached(const unsigned long long *p, unsigned long long target,
                                                   const WaitPolicy &wp, unsigned int code)
{
    if ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) >= target) return true;
    const long long t0 = clock64();
    unsigned int polls = 0;
    for (;;) {
        if ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) >= target) return true;
        if ((++polls & 63u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int *>(wp.status) != 0u) return false;
            if (wp.timeout_cycles > 0 && clock64() - t0 > wp.timeout_cycles) {
                if (atomicCAS(wp.status, 0u, code) == 0u && wp.host_status != nullptr) {
                    *reinterpret_cast<volatile unsigned int *>(wp.host_status) = code;
                    __threadfence_system();
                }
                return false;
            }
        }
        __nanosleep(SYS ? 64 : 20);
    }
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One thread: every peer has published `step`, i.e. its positions of that step
// have landed in this device's buffer.  All flags are polled together (the loads
// are independent, one round trip for all peers) and one system-scope fence makes
// the poll an acquire; the tiles are then fetched by the async proxy (bulk
// copies), hence the proxy fence.
__device__ __forceinline__ bool wait_for_peers_one_thread(const PeerExchange &X, unsigned long long step,
                                                          const WaitPolicy &wp)
{
    const long long t0 = clock64();
    unsigned int polls = 0;
    bool ok = true;
    for (;;) {
        unsigned long long lowest = ~0ull;
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            if (r < X.nranks && r != X.rank) lowest = min(lowest, ld_relaxed_sys(X.flags + r));
        }
        if (lowest >= step) break;
        if ((++polls & 63u) == 0u) {
            if (*reinterpret_cast<volatile unsigned int *>(wp.status) != 0u) { ok = false; break; }
            if (wp.timeout_cycles > 0 && clock64() - t0 > wp.timeout_cycles) {
                if (atomicCAS(wp.status, 0u, kStatusPeerTimeout) == 0u && wp.host_status != nullptr) {
                    *reinterpret_cast<volatile unsigned int *>(wp.host_status) = kStatusPeerTimeout;
                    __threadfence_system();
                }
                ok = false;
                break;
            }
        }
        __nanosleep(40);
    }
    __threadfence_system();
    asm volatile("fence.proxy.async.global;" ::: "memory");
    return ok;
}

// Lanes 0..nranks-1 of the calling CTA: publish this rank's step number in
// every peer's flag array, one lane per peer so that the release stores travel
// over NVLink side by side.
__device__ __forceinline__ void publish_step(const PeerExchange &X, unsigned long long step)
{
    if ((int)threadIdx.x < X.nranks && (int)threadIdx.x != X.rank) {
        st_release_sys(X.peer_flags[threadIdx.x] + X.rank, step);
    }
}

// ---------------------------------------------------------------------------
// Arithmetic
// ---------------------------------------------------------------------------

// MUFU.RSQ64H: ~20-bit reciprocal square root from the high word of r2.
__device__ __forceinline__ double rsqrt_seed(double r2)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    return y;
}

// The evaluated body's temporary position, sim_gravity.c:1186-1187:
// X + VX*DT/2 evaluated as X + ((VX*DT)/2), no contraction.
__device__ __forceinline__ double half_drift(double x, double v, double dt)
{
    return __dadd_rn(x, __dmul_rn(__dmul_rn(v, dt), 0.5));
}

// One pair, sim_gravity.c:1193-1204, as 13 FP64-pipe instructions.
//   w = m * r2^(-3/2) from the seed y0:  with e = 1 - r2*y0^2 (one exact FMA,
//   y0^2 is exact because y0 has 20 significant bits),
//   r2^(-3/2) = y0^3 * (1-e)^(-3/2) = y0^3 * (1 + e*(3/2 + 15/8 e) + O(e^3)),
//   |e| < 2^-19 so the dropped term is < 2^-56.
// r2 == 0 (an exactly coincident pair, :1199-1201) or a non-finite r2 yields a
// NaN that poisons the sum; commit_slice detects it and redoes that body on
// the guarded path, which keeps the test out of this loop.
__device__ __forceinline__ double pair_weight(double dx, double dy, double mj)
{
    const double r2 = fma(dy, dy, dx * dx);
    const double y0 = rsqrt_seed(r2);
    const double b  = y0 * y0;
    const double e  = fma(-r2, b, 1.0);
    const double a  = mj * y0;
    const double w0 = a * b;
    const double p  = fma(1.875, e, 1.5);
    const double g  = fma(e, p, 1.0);
    return w0 * g;
}

// r2^(-3/2) alone, 12 FP64-pipe instructions per pair: used for j-tiles whose
// bodies all carry the same mass m, where sum_j m*w_j*d_j = m * sum_j w_j*d_j and
// the multiplication by m moves out of the pair loop (one FMA per tile).
__device__ __forceinline__ double pair_weight_unit_mass(double dx, double dy)
{
    const double r2 = fma(dy, dy, dx * dx);
    const double y0 = rsqrt_seed(r2);
    const double b  = y0 * y0;
    const double e  = fma(-r2, b, 1.0);
    const double c  = y0 * b;
    const double p  = fma(1.875, e, 1.5);
    const double g  = fma(e, p, 1.0);
    return c * g;
}

// The same pair exactly as the reference evaluates it (sim_gravity.c:1196-1204):
// correctly rounded IEEE sqrt and divide, no FMA contraction.
__device__ __forceinline__ void pair_strict(double xj, double yj, double mj, double px, double py,
                                            double &sx, double &sy)
{
    const double dx = __dsub_rn(xj, px);
    const double dy = __dsub_rn(yj, py);
    const double r  = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    if (r == 0.0) return;                                        // :1199-1201
    const double t  = __ddiv_rn(mj, __dmul_rn(__dmul_rn(r, r), r));  // :1202
    sx = __dadd_rn(sx, __dmul_rn(t, dx));                        // :1203
    sy = __dadd_rn(sy, __dmul_rn(t, dy));                        // :1204
}

// sim_gravity.c:1207-1221 with the reference's association and no contraction:
//   A = G*SUM; D = V1*DT + ((0.5*A)*DT)*DT; V2 = V1 + A*DT; NEXT = P + D.
__device__ __forceinline__ void kick_drift(double sum, double dt, double p, double v,
                                           double &p_next, double &v_next, double &a)
{
    a = __dmul_rn(kG, sum);
    const double d = __dadd_rn(__dmul_rn(v, dt), __dmul_rn(__dmul_rn(__dmul_rn(0.5, a), dt), dt));
    v_next = __dadd_rn(v, __dmul_rn(a, dt));
    p_next = __dadd_rn(p, d);
}

// ---------------------------------------------------------------------------
// Compensated accumulation
// ---------------------------------------------------------------------------

// hi + err gains t with the rounding error of the addition kept in err
// (Knuth's two-sum, 6 additions, branch free, no contraction possible).  Used once
// per j-tile and body -- 0.03 additions per pair -- and when partial rows are
// added up, so that a large early term no longer costs every later addition an
// ulp of ITS size: the sum of N terms carries the error of one 256-term tile sum
// instead of the reference's N sequential additions.
__device__ __forceinline__ void two_sum_acc(double &hi, double &err, double t)
{
    const double s  = __dadd_rn(hi, t);
    const double bb = __dsub_rn(s, hi);
    const double e  = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(t, bb));
    err = __dadd_rn(err, e);
    hi  = s;
}

// ---------------------------------------------------------------------------
// Commit helpers: repair poisoned bodies, kick + drift, hand the result to the peers
// ---------------------------------------------------------------------------

// Guarded evaluation of one body by a whole CTA: the reference's own formula
// per pair (sqrt, divide, both skips), lanes striding over j, fixed-shape tree
// reduction (deterministic).  Only bodies whose fast sum came out non-finite
// get here -- in practice bodies that coincide exactly with another body.
__device__ inline void block_guarded_body(const double2 *pos_cur, const double *mass, int64_t n, int64_t gi,
                                          double px, double py, double *s_red, double &sx, double &sy)
{
    double tx = 0.0, ty = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        if (j == gi) continue;                            // sim_gravity.c:1191
        const double2 pj = __ldcg(pos_cur + j);
        pair_strict(pj.x, pj.y, mass[j], px, py, tx, ty);
    }
    s_red[threadIdx.x]              = tx;
    s_red[blockDim.x + threadIdx.x] = ty;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            s_red[threadIdx.x]              += s_red[threadIdx.x + w];
            s_red[blockDim.x + threadIdx.x] += s_red[blockDim.x + threadIdx.x + w];
        }
        __syncthreads();
    }
    sx = s_red[0];
    sy = s_red[blockDim.x];
    __syncthreads();
}

// Kick + drift of one body and the hand-over of its new position: to this
// device's next buffer, to every peer's (barrier B2 + the commit copy,
// sim_gravity.c:1226, :1242-1245) and, on a day change, to the PATH ring (:1248-1251).
__device__ __forceinline__ void commit_body(const TiledParams &P, const double2 *pos_cur, int nxt, int64_t li,
                                            double sx, double sy, long long ring_row)
{
    const int64_t gi = P.i_global0 + li;
    const double2 p  = __ldcg(pos_cur + gi);
    const double2 v  = __ldcg(P.vel + li);
    double nx, ny, nvx, nvy, ax, ay;
    kick_drift(sx, P.dt, p.x, v.x, nx, nvx, ax);
    kick_drift(sy, P.dt, p.y, v.y, ny, nvy, ay);
    if (P.export_accel) {
        P.accel_out[li] = make_double2(ax, ay);          // sim_gravity.c:1207-1208
    } else {
        const double2 pn = make_double2(nx, ny);
        P.pos[nxt][gi] = pn;                             // NEXT_X, NEXT_Y  (:1218-1219)
        P.vel[li]      = make_double2(nvx, nvy);         // NEXT_VX, NEXT_VY (:1220-1221)
        if (P.px.enabled) {
            for (int r = 0; r < P.px.nranks; ++r) {
                if (r != P.px.rank) P.px.peer_pos[nxt][r][gi] = pn;
            }
        }
        if (ring_row >= 0) P.clock.ring[ring_row * P.clock.npad + gi] = pn;
    }
}

// ---------------------------------------------------------------------------
// K1: tiled pair accumulation, all steps of a batch in one launch (the hot kernel)
// ---------------------------------------------------------------------------
//
// Work of a step = n_iblock x n_tile units (IB i-bodies against one j-tile).  The
// host flattens the units i-block-major and gives each CTA an equal contiguous
// run (a few Segments), so that one wave of SMs x ctas_per_sm CTAs finishes
// together whatever N is.  Each thread keeps R i-bodies (half-drifted position
// + accumulators) in registers; j-tiles arrive in a 4-deep shared-memory ring
// filled by 1-D bulk copies issued by thread 0 and signalled through mbarriers;
// every lane reads the same j, so the LDS are broadcasts.
//
// The grid is launched cooperatively (every CTA resident) and runs ALL steps of
// a gravity_cuda_step() batch.  There is no grid-wide barrier; the reference's
// two spin barriers (sim_gravity.c:1140, :1226) become data-flow counters:
//   arrivals[b]      partial rows of i-block b written (every CTA that holds a
//                    segment of b adds one per step);
//   commit           once all rows of b are there, EVERY CTA that contributed a
//                    row commits its own 1/rows slice of b's bodies (sums the
//                    rows in a fixed order, repairs poisoned bodies, kick +
//                    drift, stores to self, peers and the PATH ring) -- the
//                    reduction is spread over the CTAs instead of one straggler;
//   iblock_ready[b]  set by the last slice to the step number b has reached; the
//                    next step's segment of b and every fetch of a j-tile made of
//                    b's bodies wait for it, nothing else does;
//   peer flags       when the rank's last i-block of a step is committed the step
//                    number goes to every peer; remote j-tiles wait for them.
// Two position buffers suffice: i-block b can only be committed for step s+1 after
// every tile of step s+1 was consumed, hence after every i-block (and every peer)
// finished step s, hence after every CTA anywhere finished reading the buffer that
// commit overwrites.

template <int THREADS, int R, bool DIAG, bool FOLD>
__device__ __forceinline__ void consume_tile(const double2 *__restrict__ sp, const double *__restrict__ sm,
                                             const double (&xi)[R], const double (&yi)[R],
                                             double (&ax)[R], double (&ay)[R],
                                             const int64_t (&gi)[R], int64_t gj0)
{
    // pairs per unrolled iteration, measured (profiles/r2_unroll_ab.log): 16 for R = 1 and 2, 32 for the
    // 4-bodies-per-thread shape (same instruction counts either way; ptxas' register assignment
    // decides +-1 %)
    constexpr int kUnroll = (R == 4 ? (FOLD ? kPairsPerIterationFoldR4 : kPairsPerIterationGeneralR4) : kPairsPerIteration) / R;
#pragma unroll kUnroll
    for (int j = 0; j < kTile; ++j) {
        const double2 pj = sp[j];
        const double  mj = FOLD ? 0.0 : sm[j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double dx = pj.x - xi[r];
            const double dy = pj.y - yi[r];
            double w = FOLD ? pair_weight_unit_mass(dx, dy) : pair_weight(dx, dy, mj);
            if (DIAG) {
                // self is skipped by index (sim_gravity.c:1191): its distance
                // is |V*DT/2|, not 0, because only the i-body is drifted.
                w = (gj0 + j == gi[r]) ? 0.0 : w;
            }
            ax[r] = fma(w, dx, ax[r]);                   // sim_gravity.c:1203
            ay[r] = fma(w, dy, ay[r]);                   // sim_gravity.c:1204
        }
    }
}

// Slice `k` of `rows` of i-block `iblock`: wait until every partial row of the
// block has been written, add the rows up for this slice's bodies (lanes share a
// body when the slice is narrower than the CTA; fixed order, run-to-run
// deterministic), redo poisoned bodies on the guarded path, commit.  The last
// slice of the block marks it ready for the next step; the last block of the
// rank publishes the step to the peers.  Kept out of line so that it does not
// disturb the register allocation of the pair loop.
template <int THREADS, int R>
__device__ __noinline__ void commit_slice(const TiledParams &P, const double2 *pos_cur, int nxt, int iblock, int row,
                                          unsigned long long epoch, unsigned long long step_after, long long ring_row,
                                          double *s_red, int *s_bad, int *s_nbad, volatile int *s_abort)
{
    constexpr int IB  = THREADS * R;
    const int     tid = threadIdx.x;
    const int     row0 = P.iblock_seg_begin[iblock];
    const int     rows = P.iblock_seg_begin[iblock + 1] - row0;
    const int     k    = row - row0;
    const int     lo   = (int)(((long long)k * IB) / rows);
    const int     hi   = (int)(((long long)(k + 1) * IB) / rows);
    const int64_t li0  = (int64_t)iblock * IB;

    if (tid == 0) {
        if (!spin_until_reached<false>(P.arrivals + iblock, (unsigned long long)rows * (epoch + 1), P.wp,
                                       kStatusLocalTimeout))
            *s_abort = 1;
    }
    __syncthreads();
    if (*s_abort) return;

    // G lanes per body: the widest power of two that still covers the slice in one pass
    int G = 1;
    while (G < 32 && G * 2 * (hi - lo) <= THREADS && G * 2 <= rows) G <<= 1;
    const int per_pass = THREADS / G;
    const int g        = tid & (G - 1);
    for (int base = lo; base < hi; base += per_pass) {
        const int  bi    = base + tid / G;                 // body inside the i-block
        const bool valid = bi < hi && li0 + bi < P.n_local;
        double sxh = 0.0, sxl = 0.0, syh = 0.0, syl = 0.0;
        if (valid) {
            const double2 *ph = P.partial_hi + (size_t)row0 * IB + bi;
            const double2 *pl = P.partial_lo + (size_t)row0 * IB + bi;
            int r = g;
            for (; r + 3 * G < rows; r += 4 * G) {          // four rows in flight
                const double2 h0 = __ldcg(ph + (size_t)r * IB), h1 = __ldcg(ph + (size_t)(r + G) * IB);
                const double2 h2 = __ldcg(ph + (size_t)(r + 2 * G) * IB), h3 = __ldcg(ph + (size_t)(r + 3 * G) * IB);
                const double2 l0 = __ldcg(pl + (size_t)r * IB), l1 = __ldcg(pl + (size_t)(r + G) * IB);
                const double2 l2 = __ldcg(pl + (size_t)(r + 2 * G) * IB), l3 = __ldcg(pl + (size_t)(r + 3 * G) * IB);
                two_sum_acc(sxh, sxl, h0.x); two_sum_acc(syh, syl, h0.y); sxl += l0.x; syl += l0.y;
                two_sum_acc(sxh, sxl, h1.x); two_sum_acc(syh, syl, h1.y); sxl += l1.x; syl += l1.y;
                two_sum_acc(sxh, sxl, h2.x); two_sum_acc(syh, syl, h2.y); sxl += l2.x; syl += l2.y;
                two_sum_acc(sxh, sxl, h3.x); two_sum_acc(syh, syl, h3.y); sxl += l3.x; syl += l3.y;
            }
            for (; r < rows; r += G) {
                const double2 h0 = __ldcg(ph + (size_t)r * IB);
                const double2 l0 = __ldcg(pl + (size_t)r * IB);
                two_sum_acc(sxh, sxl, h0.x); two_sum_acc(syh, syl, h0.y); sxl += l0.x; syl += l0.y;
            }
        }
        // combine the G lanes of a body: butterfly, both partners compute the same sum
        for (int off = 1; off < G; off <<= 1) {
            const double oxh = __shfl_xor_sync(0xffffffffu, sxh, off), oxl = __shfl_xor_sync(0xffffffffu, sxl, off);
            const double oyh = __shfl_xor_sync(0xffffffffu, syh, off), oyl = __shfl_xor_sync(0xffffffffu, syl, off);
            // order the operands by lane so that both partners round identically
            if (g & off) {
                double th = oxh, tl = oxl; two_sum_acc(th, tl, sxh); sxh = th; sxl = tl + sxl;
                th = oyh; tl = oyl;        two_sum_acc(th, tl, syh); syh = th; syl = tl + syl;
            } else {
                two_sum_acc(sxh, sxl, oxh); sxl += oxl;
                two_sum_acc(syh, syl, oyh); syl += oyl;
            }
        }
        double sx = sxh + sxl, sy = syh + syl;
        const bool owner = valid && g == 0;
        const bool bad   = owner && !(isfinite(sx) && isfinite(sy));
        if (__syncthreads_or(bad)) {
            // poisoned bodies of this pass, one after the other, each by the whole CTA.
            // Every position of this step is complete: all rows of the block have
            // arrived, so every tile -- local and remote -- was fetched after its
            // producer had published it.
            if (tid == 0) *s_nbad = 0;
            __syncthreads();
            if (bad) s_bad[atomicAdd(s_nbad, 1)] = tid;
            __syncthreads();
            const int nbad = *s_nbad;
            for (int b = 0; b < nbad; ++b) {
                const int     owner_tid = s_bad[b];
                const int64_t oli       = li0 + base + owner_tid / G;
                const int64_t ogi       = P.i_global0 + oli;
                const double2 p         = __ldcg(pos_cur + ogi);
                const double2 v         = __ldcg(P.vel + oli);
                double fx, fy;
                block_guarded_body(pos_cur, P.mass, P.n, ogi, half_drift(p.x, v.x, P.dt), half_drift(p.y, v.y, P.dt),
                                   s_red, fx, fy);
                if (tid == owner_tid) {
                    sx = fx;
                    sy = fy;
                }
            }
        }
        if (owner) commit_body(P, pos_cur, nxt, li0 + bi, sx, sy, ring_row);
    }

    // ---- this slice is done; the last one marks the block ready for the next step ----
    // Only what the NEXT STEP ON THIS DEVICE needs happens here: a gpu-scope fence and the
    // block's counters.  Making the stores to the peers visible costs a system-scope fence
    // (an NVLink round trip); that is done once per CTA and step, off this path (peer_handover).
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned long long slices = atomicAdd(P.slices_done + iblock, 1ull) + 1ull;
        if (slices == (unsigned long long)rows * (epoch + 1) && !P.export_accel) {
            __threadfence();
            st_release_gpu(P.iblock_ready + iblock, step_after);
        }
    }
}

// After a CTA has committed all its slices of a step: warp 1 makes the CTA's stores to the
// peers visible system-wide (ONE system-scope fence per CTA and step, while thread 0 already
// waits for the next step's i-blocks) and counts the CTA's slices; the rank's last CTA of the
// step publishes the step number to every peer, one lane per peer.  The peers need that only
// when they reach this rank's tiles, which every CTA schedules after its own shard's.
template <int THREADS>
__device__ __noinline__ void peer_handover(const TiledParams &P, int n_slices, unsigned long long step_after)
{
    const int tid = threadIdx.x;
    if ((tid >> 5) != 1) return;
    const int lane = tid & 31;
    unsigned long long sent = 0;
    if (lane == 0) {
        __threadfence_system();
        sent = atomicAdd(P.xfer_done, (unsigned long long)n_slices) + (unsigned long long)n_slices;
    }
    sent = __shfl_sync(0xffffffffu, sent, 0);
    if (sent == P.xfer_base + (unsigned long long)P.n_seg_total * (step_after - P.step_base)) {
        if (lane == 0) __threadfence_system();       // every other CTA's count was preceded by its fence
        __syncwarp();
        if (lane < P.px.nranks && lane != P.px.rank) st_release_sys(P.px.peer_flags[lane] + P.px.rank, step_after);
    }
}

#ifndef GRAVITY_TILE_TWOSUM
#define GRAVITY_TILE_TWOSUM 1      // developer switch: 0 = general-mass tiles accumulate straight into the running sum
#endif
#ifdef GRAVITY_MAXNREG
#define GRAVITY_KERNEL_BOUNDS(T, M) __maxnreg__(GRAVITY_MAXNREG)
#else
#define GRAVITY_KERNEL_BOUNDS(T, M) __launch_bounds__(T, M)
#endif

// FOLDABLE = false compiles the equal-mass tile path out (option mass_fold = 0, or no
// tile of the system qualifies): the general-mass pair loop then gets the register
// allocation to itself.
template <int THREADS, int R, int MINB, bool FOLDABLE>
__global__ void GRAVITY_KERNEL_BOUNDS(THREADS, MINB) k_tiled_steps(const TiledParams P)
{
    constexpr int      IB         = THREADS * R;
    constexpr int      kWarps     = THREADS / 32;
    constexpr uint32_t kTileBytes = kTile * (sizeof(double2) + sizeof(double));

    __shared__ alignas(128) double2 s_pos[kStages][kTile];
    __shared__ alignas(128) double  s_mass[kStages][kTile];
    __shared__ alignas(8) uint64_t  s_full[kStages];
    __shared__ alignas(8) uint64_t  s_empty[kStages];
    __shared__ double               s_red[2 * THREADS];   // guarded path only
    __shared__ int                  s_bad[THREADS];
    __shared__ int                  s_nbad;
    __shared__ volatile int         s_abort;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;

    const int seg_begin = P.cta_seg_begin[blockIdx.x];
    const int seg_end   = P.cta_seg_begin[blockIdx.x + 1];
    unsigned long long *trace = P.trace ? P.trace + (size_t)blockIdx.x * kTraceEvents : nullptr;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kWarps);
        }
        mbar_fence_init();
        s_abort = 0;
        if (trace) trace[0] = global_timer_ns();
    }
    __syncthreads();

    StepClock clk = P.clock;
    int issued = 0, consumed = 0;            // tiles over the whole launch: the ring never restarts

    for (long long step = 0; step < P.nsteps; ++step) {
        const int cur = P.cur ^ (int)(step & 1);
        const double2 *pos_cur = P.pos[cur];
        const unsigned long long state_step = P.step_base + (unsigned long long)step;   // what pos_cur holds
        const unsigned long long epoch      = P.epoch + (unsigned long long)step;
        const long long ring_row = P.export_accel ? -1 : clock_tick(clk);

        if (step > 0) {
            // somebody gave up waiting: stop here, store nothing more
            if (tid == 0 && *reinterpret_cast<volatile unsigned int *>(P.wp.status) != 0u) s_abort = 1;
            __syncthreads();
        }
        if (s_abort) break;

        // ---- producer cursor (meaningful in thread 0 only) ----
        int  p_seg = seg_begin, p_tile = 0, p_tile_end = 0, p_ready_lo = 0, p_ready_hi = -1;
        bool peers_pending = P.px.enabled != 0;
        if (tid == 0 && p_seg < seg_end) {
            p_tile     = P.seg[p_seg].tile_begin;
            p_tile_end = P.seg[p_seg].tile_end;
        }
        auto issue_next = [&]() {
            if (p_seg >= seg_end) return;
            const int stage = issued % kStages;
            if (issued >= kStages) mbar_wait(&s_empty[stage], ((issued / kStages) - 1) & 1);
            if (p_tile >= P.local_tile_begin && p_tile < P.local_tile_end) {
                if (step > 0) {
                    // a tile of this device's own bodies: the i-blocks it is made of must have
                    // been committed for this step
                    const long long first = (long long)p_tile * kTile - P.i_global0;
                    const int b_lo = (int)(first / IB);
                    const int b_hi = min(P.n_iblock - 1, (int)((first + kTile - 1) / IB));
                    for (int b = b_lo; b <= b_hi; ++b) {
                        if (b >= p_ready_lo && b <= p_ready_hi) continue;
                        if (!spin_until_reached<false>(P.iblock_ready + b, state_step, P.wp, kStatusLocalTimeout))
                            s_abort = 1;
                    }
                    if (b_lo <= b_hi) { p_ready_lo = b_lo; p_ready_hi = b_hi; }
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
            } else if (peers_pending) {
                // first tile of other ranks' bodies in this step: their stores of the previous
                // step must have landed.  Tiles of this rank's own shard come first in every
                // CTA's tile order, so the exchange hides behind them.
                if (!wait_for_peers_one_thread(P.px, state_step, P.wp)) s_abort = 1;
                peers_pending = false;
            }
            mbar_arrive_expect_tx(&s_full[stage], kTileBytes);
            bulk_load(&s_pos[stage][0], pos_cur + (size_t)p_tile * kTile, kTile * sizeof(double2), &s_full[stage]);
            bulk_load(&s_mass[stage][0], P.mass + (size_t)p_tile * kTile, kTile * sizeof(double), &s_full[stage]);
            ++issued;
            if (++p_tile == p_tile_end) {
                if (++p_seg < seg_end) {
                    p_tile     = P.seg[p_seg].tile_begin;
                    p_tile_end = P.seg[p_seg].tile_end;
                }
            }
        };

        for (int s = seg_begin; s < seg_end; ++s) {
            const Segment sg = P.seg[s];
            const int64_t li0 = (int64_t)sg.iblock * IB;     // first local i of the block

            if (step > 0) {
                // the block's own bodies (and its partial rows) belong to the previous step
                // until its last slice has been committed
                if (tid == 0 && !spin_until_reached<false>(P.iblock_ready + sg.iblock, state_step, P.wp,
                                                           kStatusLocalTimeout))
                    s_abort = 1;
                __syncthreads();
            }
            if (tid == 0 && s == seg_begin) {
                for (int k = 0; k < kLookahead; ++k) issue_next();
            }

            double  xi[R], yi[R], ax[R], ay[R], cx[R], cy[R];
            int64_t gi[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t li = li0 + (int64_t)r * THREADS + tid;
                gi[r] = P.i_global0 + li;
                ax[r] = ay[r] = cx[r] = cy[r] = 0.0;
                if (li < P.n_local) {
                    const double2 p = __ldcg(pos_cur + gi[r]);
                    const double2 v = __ldcg(P.vel + li);
                    xi[r] = half_drift(p.x, v.x, P.dt);      // sim_gravity.c:1186
                    yi[r] = half_drift(p.y, v.y, P.dt);      // sim_gravity.c:1187
                } else {
                    xi[r] = 0.0;
                    yi[r] = 0.0;
                    gi[r] = -1;
                }
            }
            const int64_t gi_lo = P.i_global0 + li0;
            const int64_t gi_hi = gi_lo + IB;

            for (int t = sg.tile_begin; t < sg.tile_end; ++t) {
                if (tid == 0) issue_next();
                const double tile_m = FOLDABLE ? __ldg(P.tile_mass + t) : __longlong_as_double(0x7ff8000000000000LL);
                const int stage = consumed % kStages;
                mbar_wait(&s_full[stage], (consumed / kStages) & 1);
                if (trace && tid == 0 && consumed == 0) trace[1] = global_timer_ns();

                const int64_t gj0  = (int64_t)t * kTile;
                const bool    diag = gj0 < gi_hi && gj0 + kTile > gi_lo;
                const bool    fold = FOLDABLE && tile_m == tile_m;
                // every tile is summed into fresh registers and joins the running sum
                // through a two-sum
                double tx[R], ty[R];
#pragma unroll
                for (int r = 0; r < R; ++r) tx[r] = ty[r] = 0.0;
                if (fold) {
                    // all 256 masses of this tile are identical: sum without the mass,
                    // fold it in once per tile
                    if (diag) consume_tile<THREADS, R, true, true>(s_pos[stage], s_mass[stage], xi, yi, tx, ty, gi, gj0);
                    else      consume_tile<THREADS, R, false, true>(s_pos[stage], s_mass[stage], xi, yi, tx, ty, gi, gj0);
#if GRAVITY_TILE_TWOSUM
                } else if (diag) {
                    consume_tile<THREADS, R, true, false>(s_pos[stage], s_mass[stage], xi, yi, tx, ty, gi, gj0);
                } else {
                    consume_tile<THREADS, R, false, false>(s_pos[stage], s_mass[stage], xi, yi, tx, ty, gi, gj0);
                }
#else
                } else if (diag) {
                    consume_tile<THREADS, R, true, false>(s_pos[stage], s_mass[stage], xi, yi, ax, ay, gi, gj0);
                } else {
                    consume_tile<THREADS, R, false, false>(s_pos[stage], s_mass[stage], xi, yi, ax, ay, gi, gj0);
                }
#endif
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[stage]);
                ++consumed;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    two_sum_acc(ax[r], cx[r], fold ? tile_m * tx[r] : tx[r]);
                    two_sum_acc(ay[r], cy[r], fold ? tile_m * ty[r] : ty[r]);
                }
            }

            double2 *out_hi = P.partial_hi + (size_t)sg.slot * IB;
            double2 *out_lo = P.partial_lo + (size_t)sg.slot * IB;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double hx = ax[r] + cx[r], hy = ay[r] + cy[r];
                out_hi[r * THREADS + tid] = make_double2(hx, hy);
                out_lo[r * THREADS + tid] = make_double2((ax[r] - hx) + cx[r], (ay[r] - hy) + cy[r]);
            }
            // one more row of this i-block is there (the CTA barrier orders every thread's
            // row before thread 0's fence and arrival; one fence per CTA, not per thread)
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(P.arrivals + sg.iblock, 1ull);
            }
        }
        if (trace && tid == 0 && step == 0) trace[2] = global_timer_ns();

        // ---- commit duty: this CTA's slice of every i-block it contributed a row to ----
        for (int s = seg_begin; s < seg_end; ++s) {
            const Segment sg = P.seg[s];
            commit_slice<THREADS, R>(P, pos_cur, cur ^ 1, sg.iblock, sg.slot, epoch, state_step + 1, ring_row, s_red,
                                     s_bad, &s_nbad, &s_abort);
            if (s_abort) break;
        }
        if (P.px.enabled && !P.export_accel && !s_abort) {
            // the slices' stores were ordered before this point by the CTA barrier at the end of commit_slice
            __syncthreads();
            peer_handover<THREADS>(P, seg_end - seg_begin, state_step + 1);
        }
        if (trace && tid == 0 && step == 0) trace[3] = global_timer_ns();
    }
    if (trace && tid == 0) trace[4] = global_timer_ns();
}

// An empty shard still tells its peers that it has nothing to send.
__global__ void k_publish_only(const PeerExchange X, unsigned long long step)
{
    if (blockIdx.x == 0) publish_step(X, step);
}

// ---------------------------------------------------------------------------
// STRICT kernels: bit-identical to the reference loop
// ---------------------------------------------------------------------------

// Any N, any number of GPUs, one launch per step.  A CTA owns up to kStrictMaxBodies
// consecutive i-bodies.  The N pair terms tmp*XDIST, tmp*YDIST of a body (sim_gravity.c:
// 1202-1204) do not depend on one another, so warps 1..7 evaluate them side by side -- one
// j-body per thread and chunk, IEEE sqrt and divide, the latency chains that make a serial
// j loop slow running in parallel -- into a shared-memory buffer, while warp 0 adds the
// PREVIOUS chunk's terms to each body's sum in ascending j, one lane per body, exactly as
// the reference adds them.  A skipped pair (:1191, :1199-1201) is stored as +0.0: adding it
// leaves every sum a reference run can reach unchanged (a sum that starts at +0.0 never
// becomes -0.0).  Bit-identical to the reference; throughput set by the term arithmetic,
// not by the number of bodies a shard happens to have.
constexpr int kStrictThreads   = 256;
constexpr int kStrictChunk     = kStrictThreads - 32;   // j-bodies per chunk: one per evaluating thread
constexpr int kStrictRow       = kStrictChunk + 1;      // padded row of the term buffer (bank spread)
constexpr int kStrictMaxBodies = 8;                     // i-bodies per CTA

__host__ __device__ inline size_t strict_smem_bytes(int bodies_per_cta)
{
    return (size_t)2 * bodies_per_cta * kStrictRow * sizeof(double2);
}

__global__ void __launch_bounds__(kStrictThreads) k_strict_terms(const CommitParams P, int bodies_per_cta)
{
    extern __shared__ __align__(16) unsigned char strict_smem[];
    __shared__ double2 s_drift[kStrictMaxBodies];
    __shared__ int     s_flag;
    double2 *s_term = reinterpret_cast<double2 *>(strict_smem);      // [2][bodies_per_cta][kStrictRow]

    const int     tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int     B   = bodies_per_cta;
    const int64_t l0  = (int64_t)blockIdx.x * B;                      // first local body of this CTA
    const int     nb  = (int)min((int64_t)B, P.n_local - l0);
    const bool    owner = warp == 0 && lane < nb;                     // lane b owns body l0 + b
    const int64_t gi0 = P.i_global0 + l0;

    if (P.px.enabled) {
        // the peers' positions of this step must have landed before pos_cur is read
        if (tid == 0) s_flag = wait_for_peers_one_thread(P.px, P.wait_step, P.wp) ? 1 : 0;
        __syncthreads();
        if (!s_flag) return;                     // timed out: store nothing, publish nothing
    }
    double2 p = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
    if (owner) {
        p = P.pos_cur[gi0 + lane];
        v = P.vel[l0 + lane];
        s_drift[lane] = make_double2(half_drift(p.x, v.x, P.dt), half_drift(p.y, v.y, P.dt));
    }
    __syncthreads();

    const int64_t nchunks = (P.n + kStrictChunk - 1) / kStrictChunk;
    double sx = 0.0, sy = 0.0;
    // the evaluating threads keep the next chunk's j-body in flight while they work on this one
    const int t = tid - 32;
    double2 pj = make_double2(0.0, 0.0);
    double  mj = 0.0;
    if (warp >= 1 && t < P.n) {
        pj = P.pos_cur[t];
        mj = P.mass[t];
    }
    for (int64_t c = 0; c <= nchunks; ++c) {
        if (warp >= 1) {
            if (c < nchunks) {
                const int64_t j = c * kStrictChunk + t;
                const double2 cur_p = pj;
                const double  cur_m = mj;
                const int64_t jn = j + kStrictChunk;
                if (jn < P.n) {
                    pj = P.pos_cur[jn];
                    mj = P.mass[jn];
                }
                if (j < P.n) {
                    double2 *dst = s_term + (size_t)(c & 1) * B * kStrictRow + t;
                    for (int b = 0; b < nb; ++b) {
                        const double2 d  = s_drift[b];
                        const double  dx = __dsub_rn(cur_p.x, d.x);
                        const double  dy = __dsub_rn(cur_p.y, d.y);
                        const double  r  = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                        double2 term = make_double2(0.0, 0.0);
                        if (j != gi0 + b && r != 0.0) {                  // sim_gravity.c:1191, :1199-1201
                            const double w = __ddiv_rn(cur_m, __dmul_rn(__dmul_rn(r, r), r));   // :1202
                            term = make_double2(__dmul_rn(w, dx), __dmul_rn(w, dy));
                        }
                        dst[(size_t)b * kStrictRow] = term;
                    }
                }
            }
        } else if (owner && c >= 1) {
            const int64_t j0  = (c - 1) * kStrictChunk;
            const int     cnt = (int)min((int64_t)kStrictChunk, P.n - j0);
            const double2 *row = s_term + ((size_t)((c - 1) & 1) * B + lane) * kStrictRow;
#pragma unroll 4
            for (int jj = 0; jj < cnt; ++jj) {
                sx = __dadd_rn(sx, row[jj].x);                           // :1203
                sy = __dadd_rn(sy, row[jj].y);                           // :1204
            }
        }
        __syncthreads();
    }
    if (owner) {
        const int64_t li = l0 + lane, gi = gi0 + lane;
        double nx, ny, nvx, nvy, ax, ay;
        kick_drift(sx, P.dt, p.x, v.x, nx, nvx, ax);
        kick_drift(sy, P.dt, p.y, v.y, ny, nvy, ay);
        if (P.export_accel) {
            P.accel_out[li] = make_double2(ax, ay);          // sim_gravity.c:1207-1208
        } else {
            const double2 pn = make_double2(nx, ny);
            P.pos_next[gi] = pn;                             // NEXT_X, NEXT_Y  (:1218-1219)
            P.vel[li]      = make_double2(nvx, nvy);         // NEXT_VX, NEXT_VY (:1220-1221)
            if (P.px.enabled) {
                for (int r = 0; r < P.px.nranks; ++r) {
                    if (r != P.px.rank) P.px.peer_pos[P.cur_next][r][gi] = pn;
                }
            }
            if (P.ring_row) P.ring_row[gi] = pn;             // PATH sample on a day change (:1248-1251)
        }
    }
    if (P.px.enabled && !P.export_accel) {
        // when the last CTA gets here all new positions of this rank are visible
        // system-wide and the step number goes to every peer
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(P.commit_counter, 1u);
            s_flag = (prev == gridDim.x - 1) ? 1 : 0;
            if (s_flag) *P.commit_counter = 0u;              // ready for the next launch
            __threadfence_system();
        }
        __syncthreads();
        if (s_flag) publish_step(P.px, P.publish_step);
    }
}

// Large shards (more than kStrictRowsFrom bodies on a device): one thread per local i-body, j
// ascending through shared-memory tiles.  With that many bodies every SM is full of
// independent sqrt/divide chains and the serial j loop costs nothing; below, the
// term-parallel kernel above is 1.5-11x faster (measured crossover ~65K bodies).
constexpr int64_t kStrictRowsFrom = 65536;
constexpr int kStrictRowThreads = 128;

__global__ void __launch_bounds__(kStrictRowThreads) k_strict_rows(const CommitParams P)
{
    __shared__ double2 s_pos[kStrictRowThreads];
    __shared__ double  s_mass[kStrictRowThreads];
    __shared__ int     s_flag;

    const int64_t li     = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool    active = li < P.n_local;
    const int64_t gi     = P.i_global0 + li;
    double px = 0.0, py = 0.0, sx = 0.0, sy = 0.0;
    if (P.px.enabled) {
        // the peers' positions of this step must have landed before pos_cur is read
        if (threadIdx.x == 0) s_flag = wait_for_peers_one_thread(P.px, P.wait_step, P.wp) ? 1 : 0;
        __syncthreads();
        if (!s_flag) return;                     // timed out: store nothing, publish nothing
    }
    double2 p = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
    if (active) {
        p  = P.pos_cur[gi];
        v  = P.vel[li];
        px = half_drift(p.x, v.x, P.dt);
        py = half_drift(p.y, v.y, P.dt);
    }
    for (int64_t j0 = 0; j0 < P.n; j0 += kStrictRowThreads) {
        const int64_t j = j0 + threadIdx.x;
        __syncthreads();
        if (j < P.n) {
            s_pos[threadIdx.x]  = P.pos_cur[j];
            s_mass[threadIdx.x] = P.mass[j];
        }
        __syncthreads();
        const int cnt = (int)min((int64_t)kStrictRowThreads, P.n - j0);
        if (active) {
            for (int jj = 0; jj < cnt; ++jj) {
                if (j0 + jj == gi) continue;             // sim_gravity.c:1191
                pair_strict(s_pos[jj].x, s_pos[jj].y, s_mass[jj], px, py, sx, sy);
            }
        }
    }
    if (active) {
        double nx, ny, nvx, nvy, ax, ay;
        kick_drift(sx, P.dt, p.x, v.x, nx, nvx, ax);
        kick_drift(sy, P.dt, p.y, v.y, ny, nvy, ay);
        if (P.export_accel) {
            P.accel_out[li] = make_double2(ax, ay);          // sim_gravity.c:1207-1208
        } else {
            const double2 pn = make_double2(nx, ny);
            P.pos_next[gi] = pn;                             // NEXT_X, NEXT_Y  (:1218-1219)
            P.vel[li]      = make_double2(nvx, nvy);         // NEXT_VX, NEXT_VY (:1220-1221)
            if (P.px.enabled) {
                for (int r = 0; r < P.px.nranks; ++r) {
                    if (r != P.px.rank) P.px.peer_pos[P.cur_next][r][gi] = pn;
                }
            }
            if (P.ring_row) P.ring_row[gi] = pn;             // PATH sample on a day change (:1248-1251)
        }
    }
    if (P.px.enabled && !P.export_accel) {
        // when the last CTA gets here all new positions of this rank are visible
        // system-wide and the step number goes to every peer
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(P.commit_counter, 1u);
            s_flag = (prev == gridDim.x - 1) ? 1 : 0;
            if (s_flag) *P.commit_counter = 0u;              // ready for the next launch
            __threadfence_system();
        }
        __syncthreads();
        if (s_flag) publish_step(P.px, P.publish_step);
    }
}

// Stream-ordered wait for the peers' positions (download, energy, destroy): one thread.
__global__ void k_wait_peers(const PeerExchange X, unsigned long long step, const WaitPolicy wp)
{
    if (threadIdx.x == 0) wait_for_peers_one_thread(X, step, wp);
}

// N <= 1024, one device: the whole system lives in one CTA and `nsteps` steps
// run inside one launch (two __syncthreads per step replace the reference's
// two spin barriers, sim_gravity.c:1140 and :1226).
__global__ void __launch_bounds__(kSmallMax) k_small_strict_steps(double2 *pos, double2 *vel,
                                                                  const double *mass, int n, double dt,
                                                                  long long nsteps, StepClock clk)
{
    __shared__ double2 s_pos[kSmallMax];
    __shared__ double  s_mass[kSmallMax];

    const int  i      = threadIdx.x;
    const bool active = i < n;
    double2 p = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
    if (active) {
        p         = pos[i];
        v         = vel[i];
        s_mass[i] = mass[i];
    }
    for (long long s = 0; s < nsteps; ++s) {
        if (active) s_pos[i] = p;
        __syncthreads();
        if (active) {
            const double px = half_drift(p.x, v.x, dt);
            const double py = half_drift(p.y, v.y, dt);
            double sx = 0.0, sy = 0.0;
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                pair_strict(s_pos[j].x, s_pos[j].y, s_mass[j], px, py, sx, sy);
            }
            double ax, ay;
            kick_drift(sx, dt, p.x, v.x, p.x, v.x, ax);
            kick_drift(sy, dt, p.y, v.y, p.y, v.y, ay);
        }
        const long long ring_row = clock_tick(clk);          // commit block, sim_gravity.c:1232-1260
        if (ring_row >= 0 && active) clk.ring[ring_row * clk.npad + i] = p;
        __syncthreads();
    }
    if (active) {
        pos[i] = p;
        vel[i] = v;
    }
}

// N <= 32: one thread per ORDERED PAIR.  The pair terms tmp*XDIST, tmp*YDIST
// (sim_gravity.c:1202-1204) are independent of one another, so all N*(N-1) of
// them -- each a dependent sqrt -> multiply -> divide chain, the latency that
// dominates tiny systems -- are evaluated at once; thread (i,0) then adds its row
// in ascending j exactly as the reference does, so the bits are unchanged while a
// step costs one pair latency plus N additions instead of N pair latencies.
constexpr int kPairMax = 32;

__global__ void __launch_bounds__(kPairMax * kPairMax) k_small_strict_steps_pairs(double2 *pos, double2 *vel,
                                                                                const double *mass, int n,
                                                                                double dt, long long nsteps,
                                                                                StepClock clk)
{
    __shared__ double2 s_pos[kPairMax];
    __shared__ double2 s_drift[kPairMax];
    __shared__ double  s_mass[kPairMax];
    __shared__ double2 s_term[kPairMax][kPairMax + 1];
    __shared__ unsigned s_skip[kPairMax];          // bit j of s_skip[i]: pair (i,j) is skipped

    const int  i      = threadIdx.x / n;
    const int  j      = threadIdx.x - i * n;
    const bool owner  = j == 0;                    // thread (i,0) owns body i
    const bool active = i < n;
    double2 p = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
    if (active && owner) {
        p         = pos[i];
        v         = vel[i];
        s_mass[i] = mass[i];
    }
    for (long long s = 0; s < nsteps; ++s) {
        if (active && owner) {
            s_pos[i]   = p;
            s_drift[i] = make_double2(half_drift(p.x, v.x, dt), half_drift(p.y, v.y, dt));
            s_skip[i]  = 0u;
        }
        __syncthreads();
        if (active) {
            const double dx = __dsub_rn(s_pos[j].x, s_drift[i].x);
            const double dy = __dsub_rn(s_pos[j].y, s_drift[i].y);
            const double r  = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (j == i || r == 0.0) {              // sim_gravity.c:1191, :1199-1201
                atomicOr(&s_skip[i], 1u << j);
            } else {
                const double t = __ddiv_rn(s_mass[j], __dmul_rn(__dmul_rn(r, r), r));
                s_term[i][j]   = make_double2(__dmul_rn(t, dx), __dmul_rn(t, dy));
            }
        }
        __syncthreads();
        if (active && owner) {
            const unsigned skip = s_skip[i];
            double sx = 0.0, sy = 0.0;
            for (int k = 0; k < n; ++k) {
                if (skip & (1u << k)) continue;
                sx = __dadd_rn(sx, s_term[i][k].x);
                sy = __dadd_rn(sy, s_term[i][k].y);
            }
            double ax, ay;
            kick_drift(sx, dt, p.x, v.x, p.x, v.x, ax);
            kick_drift(sy, dt, p.y, v.y, p.y, v.y, ay);
            const long long ring_row = clock_tick(clk);      // commit block, sim_gravity.c:1232-1260
            if (ring_row >= 0) clk.ring[ring_row * clk.npad + i] = p;
        }
        // no third barrier: until the next one only owners touch shared memory,
        // each its own row, and only after it has finished reading it
    }
    if (active && owner) {
        pos[i] = p;
        vel[i] = v;
    }
}

// 32 < N <= 1024, one device: a cooperative grid of up to one CTA per SM, each
// owning B = ceil(N / SMs) bodies, runs all `nsteps` in one launch.  Per step every
// CTA (1) stages the time-t positions in shared memory, (2) evaluates the N terms
// of each of its bodies with all its threads -- the sqrt/divide latency chains run
// side by side -- (3) lets the owner thread add each row in ascending j, exactly
// the reference's summation order, and (4) writes the new positions to the other
// global buffer; one grid-wide barrier per step stands in for the reference's
// two spin barriers (sim_gravity.c:1140, :1226).  Bit-identical to the reference.
struct CoopParams {
    double2      *pos[2];
    double2      *vel;
    const double *mass;
    int           n;
    int           bodies_per_cta;
    int           cur;
    double        dt;
    long long     nsteps;
    StepClock     clock;
};

constexpr int kCoopThreads = 256;

__host__ __device__ inline size_t coop_smem_bytes(int n, int bodies_per_cta)
{
    // s_pos[n] double2 | s_mass[n] double | s_term[B][n] double2 | s_drift[B] double2 | s_skip[B][n] u8
    const size_t n_even = ((size_t)n + 1) & ~(size_t)1;      // keeps the double2 arrays 16-byte aligned
    return (size_t)n * 16 + n_even * 8 + (size_t)bodies_per_cta * n * 16 + (size_t)bodies_per_cta * 16 +
           (size_t)bodies_per_cta * n;
}

__global__ void __launch_bounds__(kCoopThreads) k_strict_coop_steps(const CoopParams P)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char coop_smem[];

    const int n = P.n, B = P.bodies_per_cta, tid = threadIdx.x;
    double2       *s_pos   = reinterpret_cast<double2 *>(coop_smem);
    double        *s_mass  = reinterpret_cast<double *>(s_pos + n);
    double2       *s_term  = reinterpret_cast<double2 *>(s_mass + ((n + 1) & ~1));
    double2       *s_drift = s_term + (size_t)B * n;
    unsigned char *s_skip  = reinterpret_cast<unsigned char *>(s_drift + B);

    const int  b0    = blockIdx.x * B;
    const int  nb    = min(B, n - b0);
    const bool owner = tid < nb;                 // thread b owns body b0 + b
    double2 p = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
    if (owner) {
        p = P.pos[P.cur][b0 + tid];
        v = P.vel[b0 + tid];
    }
    for (int j = tid; j < n; j += kCoopThreads) s_mass[j] = P.mass[j];

    int cur = P.cur;
    StepClock clk = P.clock;
    for (long long s = 0; s < P.nsteps; ++s) {
        const double2 *src = P.pos[cur];
        for (int j = tid; j < n; j += kCoopThreads) s_pos[j] = src[j];
        if (owner) s_drift[tid] = make_double2(half_drift(p.x, v.x, P.dt), half_drift(p.y, v.y, P.dt));
        __syncthreads();
        for (int b = 0; b < nb; ++b) {
            const double2 d  = s_drift[b];
            const int     gi = b0 + b;
            for (int j = tid; j < n; j += kCoopThreads) {
                const double dx = __dsub_rn(s_pos[j].x, d.x);
                const double dy = __dsub_rn(s_pos[j].y, d.y);
                const double r  = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                const bool skip = (j == gi) || (r == 0.0);   // sim_gravity.c:1191, :1199-1201
                s_skip[(size_t)b * n + j] = skip ? 1 : 0;
                if (!skip) {
                    const double t = __ddiv_rn(s_mass[j], __dmul_rn(__dmul_rn(r, r), r));
                    s_term[(size_t)b * n + j] = make_double2(__dmul_rn(t, dx), __dmul_rn(t, dy));
                }
            }
        }
        __syncthreads();
        if (owner) {
            const double2       *row  = s_term + (size_t)tid * n;
            const unsigned char *skip = s_skip + (size_t)tid * n;
            double sx = 0.0, sy = 0.0;
            for (int j = 0; j < n; ++j) {
                if (skip[j]) continue;
                sx = __dadd_rn(sx, row[j].x);
                sy = __dadd_rn(sy, row[j].y);
            }
            double ax, ay;
            kick_drift(sx, P.dt, p.x, v.x, p.x, v.x, ax);
            kick_drift(sy, P.dt, p.y, v.y, p.y, v.y, ay);
            P.pos[cur ^ 1][b0 + tid] = p;
        }
        const long long ring_row = clock_tick(clk);          // commit block, sim_gravity.c:1232-1260
        if (ring_row >= 0 && owner) clk.ring[ring_row * clk.npad + b0 + tid] = p;
        cur ^= 1;
        grid.sync();
    }
    if (owner) P.vel[b0 + tid] = v;
}

// ---------------------------------------------------------------------------
// K4: energy diagnostic (per-body terms; the host adds them up)
// ---------------------------------------------------------------------------

struct EnergyParams {
    const double2 *pos;
    const double2 *vel;     // local shard
    const double  *mass;
    double        *kin;     // local shard: 1/2 m |v|^2
    double        *pot;     // local shard: m_i * sum_{j != i} m_j / r_ij
    int64_t        n, i_global0, n_local;
};

constexpr int kEnergyThreads = 128;

__global__ void __launch_bounds__(kEnergyThreads) k_energy_terms(const EnergyParams P)
{
    __shared__ double2 s_pos[kEnergyThreads];
    __shared__ double  s_mass[kEnergyThreads];

    const int64_t li     = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool    active = li < P.n_local;
    const int64_t gi     = P.i_global0 + li;
    double2 p = make_double2(0.0, 0.0);
    if (active) p = P.pos[gi];
    double acc = 0.0;
    for (int64_t j0 = 0; j0 < P.n; j0 += kEnergyThreads) {
        const int64_t j = j0 + threadIdx.x;
        __syncthreads();
        if (j < P.n) {
            s_pos[threadIdx.x]  = P.pos[j];
            s_mass[threadIdx.x] = P.mass[j];
        }
        __syncthreads();
        const int cnt = (int)min((int64_t)kEnergyThreads, P.n - j0);
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const double dx = s_pos[jj].x - p.x;
            const double dy = s_pos[jj].y - p.y;
            const double r2 = fma(dy, dy, dx * dx);
            if (r2 > 0.0 && j0 + jj != gi) {
                const double y0 = rsqrt_seed(r2);
                const double e  = fma(-r2, y0 * y0, 1.0);
                const double y  = fma(y0 * e, fma(0.375, e, 0.5), y0);   // 1/r
                acc = fma(s_mass[jj], y, acc);
            }
        }
    }
    if (active) {
        const double2 v = P.vel[li];
        const double  m = P.mass[gi];
        P.kin[li] = 0.5 * m * (v.x * v.x + v.y * v.y);
        P.pot[li] = m * acc;
    }
}

// ---------------------------------------------------------------------------
// FP64 FMA-pipe microbenchmark (roofline denominator measured on the box)
// ---------------------------------------------------------------------------

constexpr int kPeakChains = 8;

// Each FMA reads two distinct register pairs plus an immediate.  That is the
// form that reaches the pipe's 1 warp-instruction / 2 cycles / sub-partition;
// an FMA with three distinct register operands is limited by register-file
// read bandwidth to ~1 per 3 cycles (tools/fp64_microbench.cu).
__global__ void __launch_bounds__(256) k_dfma_peak(double *sink, int iters, double a, double b,
                                                   long long *cycles)
{
    double acc[kPeakChains], x[kPeakChains];
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) {
        acc[c] = (double)(threadIdx.x + c);
        x[c]   = a + (double)(threadIdx.x * kPeakChains + c) * b;
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < kPeakChains; ++c) acc[c] = fma(acc[c], x[c], 1.5);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) s += acc[c];
    if (s == 123.456) sink[0] = s;                      // keeps the chains alive
    // the warp arbiter lets some CTAs of an SM finish long before others; the
    // longest-lived thread spans the whole kernel
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)cycles, (unsigned long long)(t1 - t0));
}

// One warp per j-tile: tile_mass[t] = m if all kTile masses of the tile have the
// same bits, NaN otherwise.
__global__ void k_tile_mass(const double *__restrict__ mass, double *tile_mass, int n_tile)
{
    const int t    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n_tile) return;
    const double   *m     = mass + (size_t)t * kTile;
    const long long first = __double_as_longlong(m[0]);
    bool same = true;
    for (int k = lane; k < kTile; k += 32) same = same && (__double_as_longlong(m[k]) == first);
    same = __all_sync(0xffffffffu, same);
    if (lane == 0) tile_mass[t] = same ? m[0] : __longlong_as_double(0x7ff8000000000000LL);
}

// Host SoA <-> device interleaved layout (upload / download): the host arrays
// are copied as they are and (de)interleaved here, so the host never loops over
// the bodies and pinned host buffers stream at full PCIe speed.
__global__ void k_pack_pairs(const double *__restrict__ a, const double *__restrict__ b, double2 *dst, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_double2(a[i], b[i]);
}

__global__ void k_unpack_pairs(const double2 *__restrict__ src, double *a, double *b, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double2 v = src[i];
        a[i] = v.x;
        b[i] = v.y;
    }
}

// The same for page-locked host arrays, which the device reads and writes in place over
// PCIe (no staging copies, one launch): the path of callers that keep their SoA state in
// pinned memory, where the fixed costs of five small copies would otherwise dominate a
// step of a few hundred microseconds.
__global__ void k_upload_shard_pinned(const double *__restrict__ x, const double *__restrict__ y,
                                      const double *__restrict__ vx, const double *__restrict__ vy,
                                      const double *__restrict__ mass, double2 *pos, double2 *vel, double *mass_dev,
                                      int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        pos[i] = make_double2(x[i], y[i]);
        vel[i] = make_double2(vx[i], vy[i]);
        if (mass != nullptr) mass_dev[i] = mass[i];
    }
}

__global__ void k_download_shard_pinned(const double2 *__restrict__ pos, const double2 *__restrict__ vel,
                                        double *x, double *y, double *vx, double *vy, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (x != nullptr || y != nullptr) {
            const double2 p = pos[i];
            if (x != nullptr) x[i] = p.x;
            if (y != nullptr) y[i] = p.y;
        }
        if (vx != nullptr || vy != nullptr) {
            const double2 v = vel[i];
            if (vx != nullptr) vx[i] = v.x;
            if (vy != nullptr) vy[i] = v.y;
        }
    }
}

// Parks padding bodies (global index >= n): mass 0, far away, never updated.
__global__ void k_init_padding(double2 *pos0, double2 *pos1, double *mass, int64_t n, int64_t npad)
{
    const int64_t i = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) {
        pos0[i] = make_double2(kPadCoord, kPadCoord);
        pos1[i] = make_double2(kPadCoord, kPadCoord);
        mass[i] = 0.0;
    }
}

}  // namespace gravity
