// gravity_cuda.cu -- host side of libgravity_cuda: the C ABI declared in
// include/gravity_cuda.h, device memory, streams, NCCL, launch planning.
//
// There is no CPU code path for the arithmetic in this file: every entry point
// that computes anything launches the sm_100a kernels of gravity_kernels.cuh
// and fails with an error when no such device is usable.
#include "gravity_cuda.h"
#include "gravity_kernels.cuh"

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

using namespace gravity;

namespace {

thread_local char g_create_error[200] = "";

struct DeviceCtx {
    int rank   = 0;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    ncclComm_t   comm = nullptr;
    cudaEvent_t  ev_start = nullptr, ev_stop = nullptr;

    double2 *pos[2] = {nullptr, nullptr};
    double  *mass = nullptr;
    double2 *vel = nullptr;
    double2 *accel = nullptr;
    double  *tile_mass = nullptr;  // [npad / kTile], see k_tile_mass
    double2 *gather_tmp = nullptr; // [npad] scratch for gathering velocities / accelerations

    // peer-memory exchange
    unsigned long long *flags = nullptr;           // [kMaxPeers], written by the peers
    unsigned int       *commit_counter = nullptr;
    unsigned int       *status = nullptr;          // kStatus*: != 0 makes the handle fail-stop
    unsigned int       *host_status = nullptr;     // mapped host copy of the verdict (read without a memcpy)
    double2            *peer_pos[2][kMaxPeers] = {};
    unsigned long long *peer_flags[kMaxPeers] = {};
    bool                ipc_opened[kMaxPeers] = {};
    double  *stage = nullptr;      // 5 x npad doubles: SoA staging for upload / download
    double  *kin = nullptr, *pot = nullptr;
    double2 *partial_hi = nullptr, *partial_lo = nullptr;
    Segment *seg = nullptr;
    int32_t *cta_seg_begin = nullptr;
    int32_t *iblock_seg_begin = nullptr;
    unsigned long long *counters = nullptr;        // arrivals | slices_done | iblock_ready | iblocks_done
    unsigned long long *trace = nullptr;           // [n_cta][kTraceEvents], option "trace"
    double2 *ring = nullptr;                       // PATH ring [path_capacity][npad]
    cudaEvent_t ev_copy = nullptr;                 // host buffers of upload() consumed
    long long clock_khz = 0;

    int64_t i0 = 0;          // global index of local body 0
    int64_t n_local = 0;     // real bodies in this shard

    // launch plan of the tiled kernel
    struct Phase { int n_cta = 0; int cta_seg_offset = 0; };
    Phase phase[1];
    int n_iblock = 0;
    int local_tile_begin = 0, local_tile_end = 0;   // j-tiles of this rank's own shard
    int n_seg = 0;
    int threads = 0, bpt = 0, minb = 0;

    // per-launch kernel timing (device 0 of the handle only)
    std::vector<cudaEvent_t> kev;
    size_t kev_used = 0;
};

}  // namespace

struct gravity_cuda {
    int64_t n = 0, chunk = 0, npad = 0;
    int nranks = 1;
    bool single_process = true;
    std::vector<DeviceCtx> ctx;      // local devices (all of them in single-process mode)
    int cur = 0;                     // which pos[] holds time t
    bool uploaded = false, planned = false;
    int opt_kernel = GRAVITY_CUDA_KERNEL_AUTO;
    int opt_threads = 0, opt_bpt = 0, opt_ctas_per_sm = 0;   // 0: chosen per problem size (build_plan)
    int opt_mass_fold = 1;
    int opt_exchange = 0;            // 0: NCCL all-gather, 1: peer-memory stores + flags
    int opt_persistent = 1;          // 1: all steps of a batch in one launch where the exchange allows it
    int opt_trace = 0;
    int opt_download_scope = 0;      // 0: full arrays on every rank, 1: only the caller's own shard
    int opt_debug_shard_of = 0;      // developer aid: plan and run like rank 0 of this many ranks, without peers
    int64_t opt_spin_timeout_ms = 10000;
    bool p2p_connected = false;
    bool mass_uploaded = false;
    bool dead = false;               // a device-side wait gave up: fail-stop
    unsigned long long step_id = 0;  // steps committed since create (absolute step number of the state)
    unsigned long long epoch = 0;    // evaluations done with the current plan of the tiled kernel
    unsigned long long xfer_steps = 0;   // of those, committed steps that stored to the peers
    int64_t path_capacity = 0;       // rows of the PATH ring (0: off)
    int64_t path_tail = 0;           // samples taken since path_enable
    int64_t launches = 0;
    bool timing = false;
    double *aos_stage = nullptr;     // page-locked 5*n doubles: SoA image of the caller's records (step_objects)
    char err[200] = "";
};

namespace {

int fail(gravity_cuda_t *h, const char *fmt, ...)
{
    char *dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 200, fmt, ap);
    va_end(ap);
    return -1;
}

#define CUDA_TRY(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail((h), "CUDA %s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NCCL_TRY(h, expr)                                                                     \
    do {                                                                                      \
        ncclResult_t r_ = (expr);                                                             \
        if (r_ != ncclSuccess)                                                                \
            return fail((h), "NCCL %s: %s (%s:%d)", #expr, ncclGetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// No C++ exception may cross the C ABI: host-side allocations (std::vector,
// new) that fail become an ordinary error return.
template <class F>
int no_throw(gravity_cuda_t *h, F &&body) noexcept
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return fail(h, "out of host memory");
    } catch (const std::exception &e) {
        return fail(h, "host error: %s", e.what());
    } catch (...) {
        return fail(h, "unknown host error");
    }
}

int check_device(gravity_cuda_t *h, int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(h, "no CUDA device: %s (libgravity_cuda has no CPU fallback)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(h, "device %d out of range (have %d)", device, count);
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(h, "device %d is sm_%d%d; libgravity_cuda is built for sm_100a only", device, prop.major,
                    prop.minor);
    return 0;
}

int effective_kernel(const gravity_cuda_t *h)
{
    if (h->opt_kernel != GRAVITY_CUDA_KERNEL_AUTO) return h->opt_kernel;
    return h->n <= kSmallMax ? GRAVITY_CUDA_KERNEL_STRICT : GRAVITY_CUDA_KERNEL_TILED;
}

// ---- kernel dispatch over (threads, bodies per thread) ----------------------

typedef void (*accum_kernel_t)(const TiledParams);

template <int T, int R, bool F>
accum_kernel_t accum_for_minb(int minb)
{
    switch (minb) {
    case 1: return k_tiled_steps<T, R, 1, F>;
    case 2: return k_tiled_steps<T, R, 2, F>;
    case 3: return k_tiled_steps<T, R, 3, F>;
    default: return k_tiled_steps<T, R, 4, F>;
    }
}

template <bool F>
accum_kernel_t pick_accum_f(int threads, int bpt, int minb)
{
    if (threads == 128) {
        if (bpt == 1) return accum_for_minb<128, 1, F>(minb);
        if (bpt == 2) return accum_for_minb<128, 2, F>(minb);
        if (bpt == 4) return accum_for_minb<128, 4, F>(minb);
    } else if (threads == 256) {
        if (bpt == 1) return accum_for_minb<256, 1, F>(minb);
        if (bpt == 2) return accum_for_minb<256, 2, F>(minb);
        if (bpt == 4) return accum_for_minb<256, 4, F>(minb);
    }
    return nullptr;
}

// foldable: the kernel variant that tests every j-tile for a common mass (option mass_fold)
accum_kernel_t pick_accum(int threads, int bpt, int minb, bool foldable = true)
{
    return foldable ? pick_accum_f<true>(threads, bpt, minb) : pick_accum_f<false>(threads, bpt, minb);
}

// ---- launch plan -------------------------------------------------------------

// Units = (i-block, j-tile) pairs, flattened i-block-major over the tile list
// `tile_runs` (pairs begin,end).  CTA c gets the units [bounds[c], bounds[c+1]);
// the resulting segments are appended to `segs`, `cta_begin[c]` marks where CTA
// c's segments start.
void plan_units(const std::vector<int> &tile_runs, const std::vector<int64_t> &bounds,
                std::vector<Segment> &segs, std::vector<int32_t> &cta_begin)
{
    int64_t tiles_per_block = 0;
    for (size_t k = 0; k + 1 < tile_runs.size(); k += 2) tiles_per_block += tile_runs[k + 1] - tile_runs[k];
    const int C = (int)bounds.size() - 1;
    for (int c = 0; c < C; ++c) {
        int64_t u = bounds[c];
        const int64_t u1 = bounds[c + 1];
        cta_begin.push_back((int32_t)segs.size());
        while (u < u1) {
            const int ib = (int)(u / tiles_per_block);
            int64_t off = u % tiles_per_block;       // offset inside this block's tile list
            size_t k = 0;                            // locate the run containing `off`
            while (off >= tile_runs[k + 1] - tile_runs[k]) {
                off -= tile_runs[k + 1] - tile_runs[k];
                k += 2;
            }
            const int t0 = tile_runs[k] + (int)off;
            const int64_t room = std::min<int64_t>(tile_runs[k + 1] - t0, u1 - u);
            Segment s;
            s.iblock = ib;
            s.tile_begin = t0;
            s.tile_end = t0 + (int)room;
            s.slot = 0;
            segs.push_back(s);
            u += room;
        }
    }
    cta_begin.push_back((int32_t)segs.size());
}

int64_t units_of(const std::vector<int> &tile_runs, int n_iblock)
{
    int64_t tiles_per_block = 0;
    for (size_t k = 0; k + 1 < tile_runs.size(); k += 2) tiles_per_block += tile_runs[k + 1] - tile_runs[k];
    return tiles_per_block * n_iblock;
}

// Everything about a launch plan that can be decided on the host without a GPU.
struct HostPlan {
    int threads = 0, bpt = 0, ctas = 0;              // CTA shape and CTAs per SM
    int n_iblock = 0, n_tile = 0, n_cta = 0;
    int local_tile_begin = 0, local_tile_end = 0;    // j-tiles of the rank's own shard
    std::vector<Segment> segs;
    std::vector<int32_t> cta_begin;                  // [n_cta + 1] into segs
    std::vector<int32_t> ib_begin;                   // [n_iblock + 1] partial rows per i-block
};

// Pick the CTA shape by how well its work units fill one wave of CTAs.  Bigger
// i-blocks amortise the j-tile loads over more pairs (measured relative pair
// rates below); smaller ones cut the quantisation loss when there are only a
// few units per CTA.  Explicit options win.
void choose_shape(int64_t n_local, int n_tile, int num_sms, int opt_threads, int opt_bpt, int opt_ctas,
                  int &threads, int &bpt, int &ctas)
{
    if (opt_threads && opt_bpt) {
        threads = opt_threads;
        bpt = opt_bpt;
        ctas = threads == 256 ? (bpt == 4 ? 1 : 2) : (bpt == 4 ? 2 : 4);
    } else {
        static const struct { int threads, bpt, ctas; double rate; } shapes[] = {
            {256, 4, 1, 1.000}, {256, 2, 2, 0.974}, {128, 2, 4, 0.955}, {128, 1, 4, 0.905}};
        double best = -1.0;
        threads = 256; bpt = 4; ctas = 1;
        for (const auto &sh : shapes) {
            const int64_t ib = sh.threads * sh.bpt;
            const int64_t W = ((n_local + ib - 1) / ib) * n_tile;
            if (W == 0) continue;
            const int64_t C = std::min<int64_t>(W, (int64_t)num_sms * sh.ctas);
            const double per = (double)W / (double)C;
            // pairs a CTA really computes include the padding of a ragged last i-block
            const double useful = (double)n_local / (double)(((n_local + ib - 1) / ib) * ib);
            const double score = sh.rate * useful * per / std::ceil(per) *
                                 std::min(1.0, (double)C * sh.threads / (num_sms * 256.0));
            // the shapes are listed biggest first; a smaller one has to promise more than 1 % to win
            // (at a tie the bigger i-block is the faster one in practice: N = 6,144, tools/explore.py)
            if (best < 0.0 || score > best * 1.01) {
                best = score;
                threads = sh.threads;
                bpt = sh.bpt;
                ctas = sh.ctas;
            }
        }
    }
    if (opt_ctas) ctas = opt_ctas;
}

// Cut the (i-block, j-tile) units of one rank into per-CTA segment lists.
void make_plan(int64_t n, int64_t n_local, int64_t i0, int64_t chunk, int nranks, int num_sms,
               int threads, int bpt, int ctas, HostPlan &out)
{
    out.threads = threads;
    out.bpt = bpt;
    out.ctas = ctas;
    const int IB = threads * bpt;
    const int n_tile = (int)((n + kTile - 1) / kTile);
    out.n_tile = n_tile;
    out.n_iblock = (int)((n_local + IB - 1) / IB);
    const int max_cta = num_sms * ctas;

    // Every CTA first works through its share of (i-block, own-shard j-tile)
    // units -- those bodies are already local when the step starts -- and only
    // then through its share of the other ranks' tiles, so that the exchange
    // of the previous step hides behind the first part.
    out.local_tile_begin = (int)std::min<int64_t>(n_tile, i0 / kTile);
    out.local_tile_end = (int)std::min<int64_t>(n_tile, (i0 + chunk) / kTile);
    std::vector<int> local_runs, remote_runs;
    if (nranks > 1) {
        local_runs = {out.local_tile_begin, out.local_tile_end};
        if (out.local_tile_begin > 0) { remote_runs.push_back(0); remote_runs.push_back(out.local_tile_begin); }
        if (out.local_tile_end < n_tile) { remote_runs.push_back(out.local_tile_end); remote_runs.push_back(n_tile); }
    } else {
        local_runs = {0, n_tile};
    }
    const int64_t Wa = units_of(local_runs, out.n_iblock), Wb = units_of(remote_runs, out.n_iblock);
    const int C = (int)std::min<int64_t>(Wa + Wb, max_cta);
    out.n_cta = C;
    // CTA c gets an equal share of ALL units (difference at most one), made of
    // its proportional share of the own-shard units plus whatever other-rank
    // units fill it up.
    std::vector<int64_t> bounds_a(C + 1, 0), bounds_b(C + 1, 0);
    for (int k = 1; k <= C; ++k) {
        bounds_a[k] = (Wa * k) / C;
        bounds_b[k] = std::min(Wb, std::max(bounds_b[k - 1], ((Wa + Wb) * k) / C - bounds_a[k]));
    }
    if (C > 0) bounds_b[C] = Wb;
    std::vector<Segment> seg_a, seg_b;
    std::vector<int32_t> cta_a, cta_b;
    if (C > 0) plan_units(local_runs, bounds_a, seg_a, cta_a);
    if (C > 0 && Wb > 0) plan_units(remote_runs, bounds_b, seg_b, cta_b);
    out.segs.clear();
    out.cta_begin.clear();
    for (int k = 0; k < C; ++k) {
        out.cta_begin.push_back((int32_t)out.segs.size());
        out.segs.insert(out.segs.end(), seg_a.begin() + cta_a[k], seg_a.begin() + cta_a[k + 1]);
        if (Wb > 0) out.segs.insert(out.segs.end(), seg_b.begin() + cta_b[k], seg_b.begin() + cta_b[k + 1]);
    }
    out.cta_begin.push_back((int32_t)out.segs.size());

    // The CTA that finishes an i-block last adds up its partial rows in a fixed
    // order.  Give every segment a partial row ("slot") such that the rows of
    // one i-block are contiguous and ordered by CTA: a stable sort of the
    // segment list on iblock.
    std::vector<Segment> &segs = out.segs;
    std::vector<int> order(segs.size());
    for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return segs[a].iblock < segs[b].iblock; });
    for (size_t k = 0; k < order.size(); ++k) segs[order[k]].slot = (int32_t)k;
    out.ib_begin.assign(out.n_iblock + 1, 0);
    for (const Segment &sg : segs) out.ib_begin[sg.iblock + 1]++;
    for (int b = 0; b < out.n_iblock; ++b) out.ib_begin[b + 1] += out.ib_begin[b];
}

int build_plan_impl(gravity_cuda_t *h)
{
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        const int n_tile = (int)((h->n + kTile - 1) / kTile);
        int threads = 0, bpt = 0, ctas = 0;
        choose_shape(c.n_local, n_tile, c.num_sms, h->opt_threads, h->opt_bpt, h->opt_ctas_per_sm, threads, bpt, ctas);
        accum_kernel_t probe = pick_accum(threads, bpt, ctas, h->opt_mass_fold != 0);
        if (!probe) return fail(h, "unsupported tiled kernel shape threads=%d bodies_per_thread=%d", threads, bpt);
        int occ = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, threads, 0));
        if (occ < 1) return fail(h, "tiled kernel does not fit on an SM");
        HostPlan plan;
        if (h->opt_debug_shard_of > 1) {
            const int64_t per = (h->n + h->opt_debug_shard_of - 1) / h->opt_debug_shard_of;
            const int64_t chunk = std::max<int64_t>(kTile, ((per + kTile - 1) / kTile) * kTile);
            make_plan(h->n, c.n_local, 0, chunk, h->opt_debug_shard_of, c.num_sms, threads, bpt, std::min(occ, ctas), plan);
        } else {
            make_plan(h->n, c.n_local, c.i0, h->chunk, h->nranks, c.num_sms, threads, bpt, std::min(occ, ctas), plan);
        }
        c.threads = plan.threads;
        c.bpt = plan.bpt;
        c.minb = plan.ctas;
        c.n_iblock = plan.n_iblock;
        c.local_tile_begin = plan.local_tile_begin;
        c.local_tile_end = plan.local_tile_end;
        c.phase[0].n_cta = plan.n_cta;
        c.phase[0].cta_seg_offset = 0;
        c.n_seg = (int)plan.segs.size();
        const int IB = c.threads * c.bpt;

        cudaFree(c.seg); cudaFree(c.cta_seg_begin); cudaFree(c.iblock_seg_begin); cudaFree(c.partial_hi);
        cudaFree(c.partial_lo); cudaFree(c.counters); cudaFree(c.trace);
        c.seg = nullptr; c.cta_seg_begin = nullptr; c.iblock_seg_begin = nullptr; c.partial_hi = nullptr;
        c.partial_lo = nullptr; c.counters = nullptr; c.trace = nullptr;
        // arrivals[n_iblock] | slices_done[n_iblock] | iblock_ready[n_iblock] | xfer_done[1]: monotonic,
        // zeroed here together with h->epoch (iblock_ready compares against absolute step numbers and
        // is only consulted for steps after the first of a launch, so 0 is a valid start)
        const size_t n_counters = (size_t)3 * std::max(1, c.n_iblock) + 1;
        CUDA_TRY(h, cudaMalloc(&c.counters, sizeof(unsigned long long) * n_counters));
        CUDA_TRY(h, cudaMemset(c.counters, 0, sizeof(unsigned long long) * n_counters));
        if (h->opt_trace) {
            const size_t n_trace = (size_t)std::max(1, plan.n_cta) * kTraceEvents;
            CUDA_TRY(h, cudaMalloc(&c.trace, sizeof(unsigned long long) * n_trace));
            CUDA_TRY(h, cudaMemset(c.trace, 0, sizeof(unsigned long long) * n_trace));
        }
        CUDA_TRY(h, cudaMalloc(&c.seg, sizeof(Segment) * std::max<size_t>(1, plan.segs.size())));
        CUDA_TRY(h, cudaMalloc(&c.cta_seg_begin, sizeof(int32_t) * std::max<size_t>(1, plan.cta_begin.size())));
        CUDA_TRY(h, cudaMalloc(&c.iblock_seg_begin, sizeof(int32_t) * plan.ib_begin.size()));
        CUDA_TRY(h, cudaMalloc(&c.partial_hi, sizeof(double2) * std::max<size_t>(1, plan.segs.size()) * IB));
        CUDA_TRY(h, cudaMalloc(&c.partial_lo, sizeof(double2) * std::max<size_t>(1, plan.segs.size()) * IB));
        CUDA_TRY(h, cudaMemcpy(c.seg, plan.segs.data(), sizeof(Segment) * plan.segs.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(c.cta_seg_begin, plan.cta_begin.data(), sizeof(int32_t) * plan.cta_begin.size(),
                               cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(c.iblock_seg_begin, plan.ib_begin.data(), sizeof(int32_t) * plan.ib_begin.size(),
                               cudaMemcpyHostToDevice));
    }
    h->epoch = 0;
    h->xfer_steps = 0;
    h->planned = true;
    return 0;
}

int build_plan(gravity_cuda_t *h)
{
    return no_throw(h, [&] { return build_plan_impl(h); });
}

// ---- timing helpers -------------------------------------------------------------

void hot_event(gravity_cuda_t *h, DeviceCtx &c)
{
    if (!h->timing || &c != &h->ctx[0]) return;
    if (c.kev_used >= 8192) return;
    if (c.kev_used == c.kev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        try {
            c.kev.push_back(e);
        } catch (...) {
            cudaEventDestroy(e);
            return;
        }
    }
    cudaEventRecord(c.kev[c.kev_used++], c.stream);
}

// ---- steps on every local device ----------------------------------------------------

bool p2p_active(const gravity_cuda_t *h) { return h->opt_exchange == 1 && h->nranks > 1; }

WaitPolicy wait_policy(const gravity_cuda_t *h, const DeviceCtx &c)
{
    WaitPolicy wp;
    wp.status = c.status;
    wp.host_status = c.host_status;
    wp.timeout_cycles = h->opt_spin_timeout_ms > 0 ? (long long)h->opt_spin_timeout_ms * c.clock_khz : 0;
    return wp;
}

PeerExchange peer_exchange(const gravity_cuda_t *h, const DeviceCtx &c, bool enabled)
{
    PeerExchange X;
    memset(&X, 0, sizeof(X));
    X.enabled = enabled ? 1 : 0;
    X.rank = c.rank;
    X.nranks = h->nranks;
    X.flags = c.flags;
    for (int r = 0; r < h->nranks && r < kMaxPeers; ++r) {
        X.peer_pos[0][r] = c.peer_pos[0][r];
        X.peer_pos[1][r] = c.peer_pos[1][r];
        X.peer_flags[r] = c.peer_flags[r];
    }
    return X;
}

// The commit block's clock for a batch that starts at `sim_time` (sim_gravity.c:1232-1236).
StepClock make_clock(const gravity_cuda_t *h, const DeviceCtx &c, int32_t dt, int64_t sim_time, int32_t last_day,
                     bool sampling)
{
    StepClock k;
    memset(&k, 0, sizeof(k));
    k.dt = dt;
    if (sampling && h->path_capacity > 0) {
        k.day = sim_time / 86400;
        k.rem = sim_time % 86400;
        k.tail = h->path_tail;
        k.capacity = h->path_capacity;
        k.ring = c.ring;
        k.npad = h->npad;
        k.last_day = last_day;
    }
    return k;
}

CommitParams commit_params(gravity_cuda_t *h, DeviceCtx &c, double dt, bool export_accel, long long ring_row)
{
    CommitParams P;
    memset(&P, 0, sizeof(P));
    P.pos_cur = c.pos[h->cur];
    P.pos_next = c.pos[h->cur ^ 1];
    P.cur_next = h->cur ^ 1;
    P.vel = c.vel;
    P.mass = c.mass;
    P.accel_out = c.accel;
    P.ring_row = (ring_row >= 0 && c.ring) ? c.ring + (size_t)ring_row * h->npad : nullptr;
    P.n = h->n;
    P.i_global0 = c.i0;
    P.n_local = c.n_local;
    P.export_accel = export_accel ? 1 : 0;
    P.dt = dt;
    P.commit_counter = c.commit_counter;
    P.wait_step = h->step_id;
    P.publish_step = h->step_id + 1;
    P.wp = wait_policy(h, c);
    P.px = peer_exchange(h, c, p2p_active(h));
    return P;
}

// One cooperative launch of the tiled kernel: `nsteps` steps (pairs, reduction, repair,
// kick + drift, hand-over, PATH samples), or one evaluation that only exports AX,AY.
int launch_tiled(gravity_cuda_t *h, DeviceCtx &c, double dt, long long nsteps, bool export_accel, const StepClock &clk)
{
    const DeviceCtx::Phase &ph = c.phase[0];
    if (ph.n_cta == 0) return 0;
    TiledParams P;
    memset(&P, 0, sizeof(P));
    P.pos[0] = c.pos[0];
    P.pos[1] = c.pos[1];
    P.cur = h->cur;
    P.mass = c.mass;
    P.vel = c.vel;
    P.accel_out = c.accel;
    P.partial_hi = c.partial_hi;
    P.partial_lo = c.partial_lo;
    P.seg = c.seg;
    P.cta_seg_begin = c.cta_seg_begin + ph.cta_seg_offset;
    P.iblock_seg_begin = c.iblock_seg_begin;
    P.tile_mass = h->opt_mass_fold ? c.tile_mass : nullptr;
    const size_t nb = (size_t)std::max(1, c.n_iblock);
    P.arrivals = c.counters;
    P.slices_done = c.counters + nb;
    P.iblock_ready = c.counters + 2 * nb;
    P.xfer_done = c.counters + 3 * nb;
    P.xfer_base = (unsigned long long)c.n_seg * h->xfer_steps;
    P.n_seg_total = c.n_seg;
    P.trace = c.trace;
    P.epoch = h->epoch;
    P.step_base = h->step_id;
    P.nsteps = nsteps;
    P.n = h->n;
    P.i_global0 = c.i0;
    P.n_local = c.n_local;
    P.n_iblock = c.n_iblock;
    P.export_accel = export_accel ? 1 : 0;
    P.local_tile_begin = c.local_tile_begin;
    P.local_tile_end = c.local_tile_end;
    P.dt = dt;
    P.clock = clk;
    P.wp = wait_policy(h, c);
    P.px = peer_exchange(h, c, p2p_active(h));
    accum_kernel_t k = pick_accum(c.threads, c.bpt, c.minb, P.tile_mass != nullptr);
    void *args[] = {&P};
    hot_event(h, c);
    // cooperative: every CTA is resident, which the in-kernel waits between CTAs rely on
    CUDA_TRY(h, cudaLaunchCooperativeKernel((const void *)k, dim3(ph.n_cta), dim3(c.threads), args, 0, c.stream));
    hot_event(h, c);
    h->launches++;
    return 0;
}

// `nsteps` iterations of sim_gravity.c:1183-1222 + the commit (:1242-1245) on all local
// devices, or (export_accel) one evaluation that writes AX,AY into ctx.accel and leaves
// the state alone.  `clk0` carries sim_time/last_day for PATH sampling (capacity 0: off).
int enqueue_steps(gravity_cuda_t *h, double dt, long long nsteps, bool export_accel, StepClock *clk0)
{
    if (h->dead) return fail(h, "the handle is dead: a device-side wait timed out earlier (see gravity_cuda.h)");
    const int kernel = effective_kernel(h);
    if (kernel == GRAVITY_CUDA_KERNEL_TILED && !h->planned) {
        if (build_plan(h) != 0) return -1;
    }
    const bool p2p = p2p_active(h);
    const bool gather = !export_accel && h->nranks > 1 && !p2p;          // NCCL all-gather after every step
    // how many steps one launch may carry: the tiled kernel runs a whole batch unless an
    // NCCL collective has to sit between the steps
    const bool batched = kernel == GRAVITY_CUDA_KERNEL_TILED && !export_accel && !gather && h->opt_persistent;
    StepClock host_clk;
    memset(&host_clk, 0, sizeof(host_clk));
    if (clk0) host_clk = *clk0;
    long long done = 0;
    while (done < nsteps) {
        const long long chunk = batched ? nsteps - done : 1;
        for (DeviceCtx &c : h->ctx) {
            CUDA_TRY(h, cudaSetDevice(c.device));
            StepClock dev_clk = host_clk;
            dev_clk.ring = c.ring;
            if (c.n_local == 0) {
                // an empty shard still has to tell its peers that it has nothing to send
                if (p2p && !export_accel) {
                    k_publish_only<<<1, 32, 0, c.stream>>>(peer_exchange(h, c, true), h->step_id + (unsigned long long)chunk);
                    h->launches++;
                    CUDA_TRY(h, cudaGetLastError());
                }
                continue;
            }
            if (kernel == GRAVITY_CUDA_KERNEL_TILED) {
                if (launch_tiled(h, c, dt, chunk, export_accel, dev_clk) != 0) return -1;
            } else {
                StepClock probe = host_clk;
                const long long ring_row = export_accel ? -1 : clock_tick(probe);
                const CommitParams P = commit_params(h, c, dt, export_accel, ring_row);
                // few bodies per CTA when the shard is small (every SM gets work), at most kStrictMaxBodies
                const int per_cta = (int)std::max<int64_t>(1, std::min<int64_t>(kStrictMaxBodies,
                                                                               (c.n_local + 3 * c.num_sms - 1) / (3 * c.num_sms)));
                const int grid = (int)((c.n_local + per_cta - 1) / per_cta);
                const size_t smem = strict_smem_bytes(per_cta);
                CUDA_TRY(h, cudaFuncSetAttribute(k_strict_terms, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)strict_smem_bytes(kStrictMaxBodies)));
                hot_event(h, c);
                if (c.n_local > kStrictRowsFrom) {
                    k_strict_rows<<<(unsigned)((c.n_local + kStrictRowThreads - 1) / kStrictRowThreads), kStrictRowThreads, 0,
                                    c.stream>>>(P);
                } else {
                    k_strict_terms<<<grid, kStrictThreads, smem, c.stream>>>(P, per_cta);
                }
                hot_event(h, c);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
        }
        if (kernel == GRAVITY_CUDA_KERNEL_TILED) {
            h->epoch += (unsigned long long)chunk;
            if (p2p && !export_accel) h->xfer_steps += (unsigned long long)chunk;
        }
        if (!export_accel) {
            for (long long k = 0; k < chunk; ++k) clock_tick(host_clk);
            if (gather) {
                // barrier B2 + the NEXT_* -> state copy (sim_gravity.c:1226, :1242-1245)
                // become one in-place all-gather of the new positions.
                NCCL_TRY(h, ncclGroupStart());
                for (DeviceCtx &c : h->ctx) {
                    double2 *next = c.pos[h->cur ^ 1];
                    NCCL_TRY(h, ncclAllGather(next + c.i0, next, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
                }
                NCCL_TRY(h, ncclGroupEnd());
            }
            h->step_id += (unsigned long long)chunk;
            h->cur ^= (int)(chunk & 1);
            if (host_clk.capacity > 0) h->path_tail = host_clk.tail;
        }
        done += chunk;
    }
    if (clk0 && !export_accel) *clk0 = host_clk;
    return 0;
}

// make the stream wait until every peer's positions of the current step have landed in
// this device's pos[cur] (peer-memory exchange only)
int enqueue_wait_peers(gravity_cuda_t *h, DeviceCtx &c)
{
    if (!p2p_active(h)) return 0;
    k_wait_peers<<<1, 32, 0, c.stream>>>(peer_exchange(h, c, true), h->step_id, wait_policy(h, c));
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

// A device-side wait that gave up leaves the state partly advanced: report it once and
// refuse further work (fail-stop; see "spin_timeout_ms" in gravity_cuda.h).
int check_status(gravity_cuda_t *h)
{
    for (DeviceCtx &c : h->ctx) {
        // written by the kernel that gave up, before it ended; every caller synchronised the stream first
        const unsigned int st = c.host_status ? *reinterpret_cast<volatile unsigned int *>(c.host_status) : 0u;
        if (st != 0) {
            h->dead = true;
            return fail(h, st == kStatusPeerTimeout
                               ? "peer-memory exchange timed out on rank %d: a peer never published its step; handle is dead"
                               : "step kernel timed out on rank %d waiting for an i-block of its own device; handle is dead",
                        c.rank);
        }
    }
    return 0;
}

int sync_all(gravity_cuda_t *h)
{
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    }
    return 0;
}

// every rank has finished everything it enqueued (used where a rank is about to
// overwrite state that a peer may still be storing into)
int rank_barrier(gravity_cuda_t *h)
{
    if (sync_all(h) != 0) return -1;
    if (h->single_process || h->nranks == 1) return 0;      // one host thread drives every device
    DeviceCtx &c = h->ctx[0];
    CUDA_TRY(h, cudaSetDevice(c.device));
    NCCL_TRY(h, ncclAllReduce(c.kin, c.kin, 1, ncclDouble, ncclSum, c.comm, c.stream));
    CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    return 0;
}

int init_ctx(gravity_cuda_t *h, DeviceCtx &c)
{
    CUDA_TRY(h, cudaSetDevice(c.device));
    CUDA_TRY(h, cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, c.device));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaEventCreate(&c.ev_start));
    CUDA_TRY(h, cudaEventCreate(&c.ev_stop));
    CUDA_TRY(h, cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
    int khz = 0;
    CUDA_TRY(h, cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c.device));
    c.clock_khz = khz;
    c.i0 = (int64_t)c.rank * h->chunk;
    c.n_local = std::max<int64_t>(0, std::min<int64_t>(h->chunk, h->n - c.i0));
    CUDA_TRY(h, cudaMalloc(&c.pos[0], sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.pos[1], sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.mass, sizeof(double) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.vel, sizeof(double2) * h->chunk));
    CUDA_TRY(h, cudaMalloc(&c.accel, sizeof(double2) * h->chunk));
    CUDA_TRY(h, cudaMalloc(&c.stage, sizeof(double) * 5 * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.tile_mass, sizeof(double) * (h->npad / kTile)));
    CUDA_TRY(h, cudaMalloc(&c.gather_tmp, sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.flags, sizeof(unsigned long long) * kMaxPeers));
    CUDA_TRY(h, cudaMalloc(&c.commit_counter, sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&c.status, sizeof(unsigned int)));
    CUDA_TRY(h, cudaHostAlloc(&c.host_status, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
    *c.host_status = 0u;
    CUDA_TRY(h, cudaMemsetAsync(c.flags, 0, sizeof(unsigned long long) * kMaxPeers, c.stream));
    CUDA_TRY(h, cudaMemsetAsync(c.commit_counter, 0, sizeof(unsigned int), c.stream));
    CUDA_TRY(h, cudaMemsetAsync(c.status, 0, sizeof(unsigned int), c.stream));
    CUDA_TRY(h, cudaMalloc(&c.kin, sizeof(double) * h->chunk));
    CUDA_TRY(h, cudaMalloc(&c.pot, sizeof(double) * h->chunk));
    CUDA_TRY(h, cudaMemsetAsync(c.vel, 0, sizeof(double2) * h->chunk, c.stream));
    const int64_t npadding = h->npad - h->n;
    if (npadding > 0) {
        k_init_padding<<<(int)((npadding + 255) / 256), 256, 0, c.stream>>>(c.pos[0], c.pos[1], c.mass, h->n, h->npad);
        CUDA_TRY(h, cudaGetLastError());
    }
    CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    return 0;
}

void set_geometry(gravity_cuda_t *h, int64_t n, int nranks)
{
    h->n = n;
    h->nranks = nranks;
    const int64_t per = (n + nranks - 1) / nranks;
    h->chunk = ((per + kTile - 1) / kTile) * kTile;      // shards are whole j-tiles
    if (h->chunk == 0) h->chunk = kTile;
    h->npad = h->chunk * nranks;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================

extern "C" {

int gravity_cuda_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return -1;
    int usable = 0;
    for (int d = 0; d < count; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) usable++;
    }
    return usable;
}

const char *gravity_cuda_last_error(const gravity_cuda_t *h) { return h ? h->err : g_create_error; }

int gravity_cuda_nccl_unique_id(void *id_out, size_t id_bytes)
{
    if (!id_out || id_bytes < sizeof(ncclUniqueId)) return fail(nullptr, "nccl id buffer too small");
    ncclUniqueId id;
    NCCL_TRY(nullptr, ncclGetUniqueId(&id));
    memset(id_out, 0, id_bytes);
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

static int create_common(gravity_cuda_t **out, int64_t n, int nranks, bool single_process,
                         int rank, int device, const void *nccl_id, size_t id_bytes)
{
    if (!out) return fail(nullptr, "out is NULL");
    *out = nullptr;
    if (n < 1) return fail(nullptr, "n must be >= 1 (got %lld)", (long long)n);
    if (nranks < 1 || nranks > kMaxPeers)
        return fail(nullptr, "number of GPUs must be 1..%d, the GPUs of one NVSwitch box (got %d)", kMaxPeers, nranks);
    gravity_cuda_t *h = new (std::nothrow) gravity_cuda();
    if (!h) return fail(nullptr, "out of host memory");
    h->single_process = single_process;
    set_geometry(h, n, nranks);

    auto bail = [&](void) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", h->err);
        gravity_cuda_destroy(h);
        return -1;
    };

    if (single_process) {
        h->ctx.resize(nranks);
        for (int r = 0; r < nranks; ++r) {
            h->ctx[r].rank = r;
            h->ctx[r].device = r;
            if (check_device(h, r) != 0) return bail();
        }
    } else {
        h->ctx.resize(1);
        h->ctx[0].rank = rank;
        h->ctx[0].device = device;
        if (check_device(h, device) != 0) return bail();
    }
    for (DeviceCtx &c : h->ctx) {
        if (init_ctx(h, c) != 0) return bail();
    }
    if (nranks > 1) {
        if (single_process) {
            std::vector<ncclComm_t> comms(nranks);
            std::vector<int> devs(nranks);
            for (int r = 0; r < nranks; ++r) devs[r] = r;
            ncclResult_t r_ = ncclCommInitAll(comms.data(), nranks, devs.data());
            if (r_ != ncclSuccess) {
                fail(h, "NCCL ncclCommInitAll: %s", ncclGetErrorString(r_));
                return bail();
            }
            for (int r = 0; r < nranks; ++r) h->ctx[r].comm = comms[r];
        } else {
            if (!nccl_id || id_bytes < sizeof(ncclUniqueId)) {
                fail(h, "rank mode with nranks > 1 needs the NCCL unique id");
                return bail();
            }
            ncclUniqueId id;
            memcpy(&id, nccl_id, sizeof(id));
            cudaSetDevice(device);
            ncclResult_t r_ = ncclCommInitRank(&h->ctx[0].comm, nranks, id, rank);
            if (r_ != ncclSuccess) {
                fail(h, "NCCL ncclCommInitRank: %s", ncclGetErrorString(r_));
                return bail();
            }
        }
    }
    *out = h;
    return 0;
}

int gravity_cuda_create(gravity_cuda_t **out, int64_t n, int ngpus)
{
    return no_throw(nullptr, [&] { return create_common(out, n, ngpus, true, 0, 0, nullptr, 0); });
}

int gravity_cuda_create_rank(gravity_cuda_t **out, int64_t n, int rank, int nranks, int device,
                             const void *nccl_id, size_t id_bytes)
{
    if (rank < 0 || rank >= nranks) return fail(nullptr, "rank %d out of range 0..%d", rank, nranks - 1);
    return no_throw(nullptr, [&] { return create_common(out, n, nranks, false, rank, device, nccl_id, id_bytes); });
}

void gravity_cuda_destroy(gravity_cuda_t *h)
{
    if (!h) return;
    // Nothing may be freed while a peer can still store into it.  With the peer-memory
    // exchange a peer's last step kernel writes this device's position buffer and flags
    // until it has published the handle's current step: wait for that (bounded by the
    // spin timeout), then for this device's own work.  In rank mode destroy is therefore
    // collective in the same sense as step: call it on every rank after the same steps.
    for (DeviceCtx &c : h->ctx) {
        cudaSetDevice(c.device);
        if (c.stream) {
            if (p2p_active(h) && h->p2p_connected && c.flags && c.status && !h->dead) {
                k_wait_peers<<<1, 32, 0, c.stream>>>(peer_exchange(h, c, true), h->step_id, wait_policy(h, c));
            }
            cudaStreamSynchronize(c.stream);
        }
    }
    cudaGetLastError();
    for (DeviceCtx &c : h->ctx) {
        cudaSetDevice(c.device);
        if (c.comm) ncclCommDestroy(c.comm);
        for (int r = 0; r < kMaxPeers; ++r) {
            if (c.ipc_opened[r]) {
                cudaIpcCloseMemHandle(c.peer_pos[0][r]);
                cudaIpcCloseMemHandle(c.peer_pos[1][r]);
                cudaIpcCloseMemHandle(c.peer_flags[r]);
            }
        }
        cudaFree(c.pos[0]); cudaFree(c.pos[1]); cudaFree(c.mass); cudaFree(c.vel); cudaFree(c.accel);
        cudaFree(c.kin); cudaFree(c.pot); cudaFree(c.stage); cudaFree(c.tile_mass); cudaFree(c.ring);
        cudaFree(c.gather_tmp); cudaFree(c.flags); cudaFree(c.commit_counter); cudaFree(c.status);
        cudaFree(c.partial_hi); cudaFree(c.partial_lo); cudaFree(c.seg); cudaFree(c.cta_seg_begin);
        cudaFree(c.iblock_seg_begin); cudaFree(c.counters); cudaFree(c.trace);
        for (cudaEvent_t e : c.kev) cudaEventDestroy(e);
        if (c.ev_start) cudaEventDestroy(c.ev_start);
        if (c.ev_stop) cudaEventDestroy(c.ev_stop);
        if (c.ev_copy) cudaEventDestroy(c.ev_copy);
        if (c.host_status) cudaFreeHost(c.host_status);
        if (&c == &h->ctx[0] && h->aos_stage) cudaFreeHost(h->aos_stage);
        if (c.stream) cudaStreamDestroy(c.stream);
    }
    delete h;
}

namespace {
struct P2PBlob {
    cudaIpcMemHandle_t pos[2];
    cudaIpcMemHandle_t flags;
    int rank;
};
static_assert(sizeof(P2PBlob) <= GRAVITY_CUDA_P2P_BLOB_BYTES, "blob too large");
}  // namespace

int gravity_cuda_p2p_export(gravity_cuda_t *h, void *blob, size_t bytes)
{
    if (!h || !blob) return fail(h, "p2p_export: NULL argument");
    if (bytes < GRAVITY_CUDA_P2P_BLOB_BYTES) return fail(h, "p2p_export: blob buffer too small");
    if (h->single_process) return fail(h, "p2p_export is for rank-mode handles; a single-process handle connects itself");
    DeviceCtx &c = h->ctx[0];
    P2PBlob b;
    memset(&b, 0, sizeof(b));
    CUDA_TRY(h, cudaSetDevice(c.device));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.pos[0], c.pos[0]));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.pos[1], c.pos[1]));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.flags, c.flags));
    b.rank = c.rank;
    memset(blob, 0, bytes);
    memcpy(blob, &b, sizeof(b));
    return 0;
}

int gravity_cuda_p2p_connect(gravity_cuda_t *h, const void *blobs, size_t bytes)
{
    if (!h) return fail(nullptr, "p2p_connect: NULL handle");
    if (h->nranks > kMaxPeers) return fail(h, "peer-memory exchange supports at most %d ranks", kMaxPeers);
    if (h->p2p_connected) return 0;
    if (h->single_process) {
        for (DeviceCtx &a : h->ctx) {
            CUDA_TRY(h, cudaSetDevice(a.device));
            for (DeviceCtx &b : h->ctx) {
                if (&a != &b) {
                    int can = 0;
                    CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, a.device, b.device));
                    if (!can) return fail(h, "device %d cannot access device %d's memory", a.device, b.device);
                    cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e != cudaSuccess) return fail(h, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
                }
                a.peer_pos[0][b.rank] = b.pos[0];
                a.peer_pos[1][b.rank] = b.pos[1];
                a.peer_flags[b.rank] = b.flags;
            }
        }
    } else {
        if (!blobs || bytes < (size_t)h->nranks * GRAVITY_CUDA_P2P_BLOB_BYTES)
            return fail(h, "p2p_connect: need %d blobs of %d bytes", h->nranks, GRAVITY_CUDA_P2P_BLOB_BYTES);
        DeviceCtx &c = h->ctx[0];
        CUDA_TRY(h, cudaSetDevice(c.device));
        for (int r = 0; r < h->nranks; ++r) {
            if (r == c.rank) {
                c.peer_pos[0][r] = c.pos[0];
                c.peer_pos[1][r] = c.pos[1];
                c.peer_flags[r] = c.flags;
                continue;
            }
            P2PBlob b;
            memcpy(&b, static_cast<const char *>(blobs) + (size_t)r * GRAVITY_CUDA_P2P_BLOB_BYTES, sizeof(b));
            if (b.rank != r) return fail(h, "p2p_connect: blob %d belongs to rank %d", r, b.rank);
            void *p0 = nullptr, *p1 = nullptr, *pf = nullptr;
            CUDA_TRY(h, cudaIpcOpenMemHandle(&p0, b.pos[0], cudaIpcMemLazyEnablePeerAccess));
            CUDA_TRY(h, cudaIpcOpenMemHandle(&p1, b.pos[1], cudaIpcMemLazyEnablePeerAccess));
            CUDA_TRY(h, cudaIpcOpenMemHandle(&pf, b.flags, cudaIpcMemLazyEnablePeerAccess));
            c.peer_pos[0][r] = static_cast<double2 *>(p0);
            c.peer_pos[1][r] = static_cast<double2 *>(p1);
            c.peer_flags[r] = static_cast<unsigned long long *>(pf);
            c.ipc_opened[r] = true;
        }
    }
    h->p2p_connected = true;
    return 0;
}

int gravity_cuda_set_option(gravity_cuda_t *h, const char *key, int64_t value)
{
    if (!h || !key) return fail(h, "set_option: NULL argument");
    if (!strcmp(key, "kernel")) {
        if (value < 0 || value > 2) return fail(h, "kernel must be 0 (auto), 1 (tiled) or 2 (strict)");
        h->opt_kernel = (int)value;
    } else if (!strcmp(key, "threads")) {
        if (value != 0 && value != 128 && value != 256) return fail(h, "threads must be 0 (auto), 128 or 256");
        h->opt_threads = (int)value;
    } else if (!strcmp(key, "bodies_per_thread")) {
        if (value != 0 && value != 1 && value != 2 && value != 4) return fail(h, "bodies_per_thread must be 0 (auto), 1, 2 or 4");
        h->opt_bpt = (int)value;
    } else if (!strcmp(key, "ctas_per_sm")) {
        if (value < 0 || value > 8) return fail(h, "ctas_per_sm must be 0 (auto) or 1..8");
        h->opt_ctas_per_sm = (int)value;
    } else if (!strcmp(key, "mass_fold")) {
        h->opt_mass_fold = value ? 1 : 0;
    } else if (!strcmp(key, "exchange")) {
        if (value != 0 && value != 1) return fail(h, "exchange must be 0 (NCCL all-gather) or 1 (peer memory)");
        if (value == 1 && h->nranks > 1 && !h->p2p_connected) {
            if (!h->single_process)
                return fail(h, "exchange=1 in rank mode needs gravity_cuda_p2p_connect() first");
            if (gravity_cuda_p2p_connect(h, nullptr, 0) != 0) return -1;
        }
        if (rank_barrier(h) != 0) return -1;
        if (value == 1 && h->nranks > 1) {
            // the flags count absolute steps; with the all-gather nobody published any, so every
            // rank writes the step all ranks are at (this call is collective) into its own array
            std::vector<unsigned long long> now(kMaxPeers, h->step_id);
            for (DeviceCtx &c : h->ctx) {
                CUDA_TRY(h, cudaSetDevice(c.device));
                CUDA_TRY(h, cudaMemcpy(c.flags, now.data(), sizeof(unsigned long long) * kMaxPeers, cudaMemcpyHostToDevice));
            }
            if (rank_barrier(h) != 0) return -1;
        }
        h->opt_exchange = (int)value;
    } else if (!strcmp(key, "debug_shard_of")) {
        // developer aid for tuning the small-shard regime on one GPU: the device updates only the
        // bodies rank 0 of `value` ranks would own, with that rank's launch plan and no exchange
        if (h->nranks != 1 || value < 1 || value > kMaxPeers) return fail(h, "debug_shard_of needs a 1-GPU handle and 1..%d", kMaxPeers);
        h->opt_debug_shard_of = (int)value;
        const int64_t per = (h->n + value - 1) / value;
        h->ctx[0].n_local = std::min<int64_t>(h->n, std::max<int64_t>(kTile, ((per + kTile - 1) / kTile) * kTile));
    } else if (!strcmp(key, "persistent")) {
        h->opt_persistent = value ? 1 : 0;
    } else if (!strcmp(key, "trace")) {
        h->opt_trace = value ? 1 : 0;
    } else if (!strcmp(key, "spin_timeout_ms")) {
        if (value < 0) return fail(h, "spin_timeout_ms must be >= 0 (0: wait for ever)");
        h->opt_spin_timeout_ms = value;
    } else if (!strcmp(key, "download_scope")) {
        if (value != 0 && value != 1) return fail(h, "download_scope must be 0 (all bodies) or 1 (own shard)");
        h->opt_download_scope = (int)value;
    } else {
        return fail(h, "unknown option '%s'", key);
    }
    h->planned = false;
    return 0;
}

// The device-side alias of a page-locked host array (cudaHostAlloc / cudaHostRegister, e.g. a
// pinned torch tensor), or NULL for ordinary pageable memory.
static double *pinned_alias(const double *p)
{
    if (p == nullptr) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? static_cast<double *>(a.devicePointer) : nullptr;
}

// Enqueues the whole upload; the host arrays are consumed once every device's ev_copy has fired.
static int upload_enqueue(gravity_cuda_t *h, const double *mass, const double *x, const double *y,
                          const double *vx, const double *vy)
{
    if (!h) return fail(nullptr, "upload: NULL handle");
    if (h->dead) return fail(h, "the handle is dead: a device-side wait timed out earlier");
    if (!x || !y || !vx || !vy) return fail(h, "upload: NULL array");
    if (!mass && !h->mass_uploaded) return fail(h, "upload: mass may only be NULL after an earlier upload set the masses");
    const int64_t np = h->npad;
    // Every device copies only its OWN shard from the host (40*N bytes cross PCIe in
    // total, whatever the number of GPUs); the devices then complete each other's
    // position and mass arrays with one in-place all-gather over NVLink.  That
    // collective also orders this upload after every peer's earlier step kernels,
    // which may still be storing into this device's buffers.
    // page-locked host arrays are read in place by one kernel per device; pageable ones go
    // through staging copies
    const double *px = pinned_alias(x), *py = pinned_alias(y), *pvx = pinned_alias(vx), *pvy = pinned_alias(vy);
    const double *pm = pinned_alias(mass);
    const bool in_place = px && py && pvx && pvy && (pm || !mass);
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        double *sx = c.stage, *sy = c.stage + np, *svx = c.stage + 2 * np, *svy = c.stage + 3 * np;
        if (in_place) {
            if (c.n_local > 0) {
                k_upload_shard_pinned<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(
                    px + c.i0, py + c.i0, pvx + c.i0, pvy + c.i0, pm ? pm + c.i0 : nullptr, c.pos[0] + c.i0, c.vel,
                    c.mass + c.i0, c.n_local);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
            CUDA_TRY(h, cudaEventRecord(c.ev_copy, c.stream));      // the host arrays are free again after this
            continue;
        }
        if (c.n_local > 0) {
            const size_t bytes = sizeof(double) * c.n_local;
            CUDA_TRY(h, cudaMemcpyAsync(sx, x + c.i0, bytes, cudaMemcpyHostToDevice, c.stream));
            CUDA_TRY(h, cudaMemcpyAsync(sy, y + c.i0, bytes, cudaMemcpyHostToDevice, c.stream));
            CUDA_TRY(h, cudaMemcpyAsync(svx, vx + c.i0, bytes, cudaMemcpyHostToDevice, c.stream));
            CUDA_TRY(h, cudaMemcpyAsync(svy, vy + c.i0, bytes, cudaMemcpyHostToDevice, c.stream));
            if (mass) CUDA_TRY(h, cudaMemcpyAsync(c.mass + c.i0, mass + c.i0, bytes, cudaMemcpyHostToDevice, c.stream));
        }
        CUDA_TRY(h, cudaEventRecord(c.ev_copy, c.stream));          // the host arrays are free again after this
        if (c.n_local > 0) {
            const unsigned grid = (unsigned)((c.n_local + 255) / 256);
            k_pack_pairs<<<grid, 256, 0, c.stream>>>(sx, sy, c.pos[0] + c.i0, c.n_local);
            k_pack_pairs<<<grid, 256, 0, c.stream>>>(svx, svy, c.vel, c.n_local);
            h->launches += 2;
            CUDA_TRY(h, cudaGetLastError());
        }
    }
    if (h->nranks > 1) {
        NCCL_TRY(h, ncclGroupStart());
        for (DeviceCtx &c : h->ctx) {
            NCCL_TRY(h, ncclAllGather(c.pos[0] + c.i0, c.pos[0], (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
            if (mass) NCCL_TRY(h, ncclAllGather(c.mass + c.i0, c.mass, (size_t)h->chunk, ncclDouble, c.comm, c.stream));
        }
        NCCL_TRY(h, ncclGroupEnd());
    }
    if (mass) {
        for (DeviceCtx &c : h->ctx) {
            CUDA_TRY(h, cudaSetDevice(c.device));
            const int n_tile = (int)(np / kTile);
            k_tile_mass<<<(n_tile + 3) / 4, 128, 0, c.stream>>>(c.mass, c.tile_mass, n_tile);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
    }
    h->cur = 0;
    if (mass) h->mass_uploaded = true;
    h->uploaded = true;
    return 0;
}

int gravity_cuda_upload(gravity_cuda_t *h, const double *mass, const double *x, const double *y,
                        const double *vx, const double *vy)
{
    if (upload_enqueue(h, mass, x, y, vx, vy) != 0) return -1;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaEventSynchronize(c.ev_copy));
    }
    return 0;
}

// positions / velocities / accelerations of the caller's own shard(s): device -> host
static int download_shards(gravity_cuda_t *h, const double2 *(*src)(gravity_cuda_t *, DeviceCtx &), bool global_index,
                           double *a, double *b, int stage_slot)
{
    const int64_t np = h->npad;
    for (DeviceCtx &c : h->ctx) {
        if (c.n_local == 0) continue;
        CUDA_TRY(h, cudaSetDevice(c.device));
        double *sa = c.stage + (size_t)stage_slot * np, *sb = sa + np;
        const double2 *from = src(h, c) + (global_index ? c.i0 : 0);
        k_unpack_pairs<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(from, sa, sb, c.n_local);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        if (a) CUDA_TRY(h, cudaMemcpyAsync(a + c.i0, sa, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
        if (b) CUDA_TRY(h, cudaMemcpyAsync(b + c.i0, sb, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
    }
    return 0;
}

static const double2 *src_pos(gravity_cuda_t *h, DeviceCtx &c) { return c.pos[h->cur]; }
static const double2 *src_vel(gravity_cuda_t *, DeviceCtx &c) { return c.vel; }
static const double2 *src_accel(gravity_cuda_t *, DeviceCtx &c) { return c.accel; }

// rank mode, all bodies wanted: gather a sharded double2 array through the scratch buffer
static int download_gathered(gravity_cuda_t *h, const double2 *shard, double *a, double *b, int stage_slot)
{
    const int64_t n = h->n, np = h->npad;
    DeviceCtx &c = h->ctx[0];
    CUDA_TRY(h, cudaSetDevice(c.device));
    double2 *tmp = c.gather_tmp;
    double *sa = c.stage + (size_t)stage_slot * np, *sb = sa + np;
    CUDA_TRY(h, cudaMemcpyAsync(tmp + c.i0, shard, sizeof(double2) * h->chunk, cudaMemcpyDeviceToDevice, c.stream));
    NCCL_TRY(h, ncclAllGather(tmp + c.i0, tmp, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
    k_unpack_pairs<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(tmp, sa, sb, n);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (a) CUDA_TRY(h, cudaMemcpyAsync(a, sa, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    if (b) CUDA_TRY(h, cudaMemcpyAsync(b, sb, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    return 0;
}

// Enqueues the whole download; the host arrays are complete after the streams have been synchronised.
static int download_enqueue(gravity_cuda_t *h, double *x, double *y, double *vx, double *vy)
{
    if (!h) return fail(nullptr, "download: NULL handle");
    if (!h->uploaded) return fail(h, "download before upload");
    const int64_t n = h->n, np = h->npad;
    // a handle that holds every shard itself, or a caller that asked for its own shard
    // only ("download_scope" = 1), needs no exchange: every device returns what it owns
    const bool own_only = h->single_process || h->nranks == 1 || h->opt_download_scope == 1;
    if (own_only) {
        // page-locked destinations are written in place by one kernel per device
        double *px = pinned_alias(x), *py = pinned_alias(y), *pvx = pinned_alias(vx), *pvy = pinned_alias(vy);
        if ((px || !x) && (py || !y) && (pvx || !vx) && (pvy || !vy)) {
            for (DeviceCtx &c : h->ctx) {
                if (c.n_local == 0) continue;
                CUDA_TRY(h, cudaSetDevice(c.device));
                k_download_shard_pinned<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(
                    c.pos[h->cur] + c.i0, c.vel, px ? px + c.i0 : nullptr, py ? py + c.i0 : nullptr,
                    pvx ? pvx + c.i0 : nullptr, pvy ? pvy + c.i0 : nullptr, c.n_local);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
            return 0;
        }
    }
    if (x || y) {
        if (own_only) {
            if (download_shards(h, src_pos, true, x, y, 0) != 0) return -1;
        } else {
            DeviceCtx &c = h->ctx[0];
            CUDA_TRY(h, cudaSetDevice(c.device));
            double *sx = c.stage, *sy = c.stage + np;
            if (enqueue_wait_peers(h, c) != 0) return -1;
            k_unpack_pairs<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(c.pos[h->cur], sx, sy, n);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            if (x) CUDA_TRY(h, cudaMemcpyAsync(x, sx, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
            if (y) CUDA_TRY(h, cudaMemcpyAsync(y, sy, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
        }
    }
    if (vx || vy) {
        if (own_only) {
            if (download_shards(h, src_vel, false, vx, vy, 2) != 0) return -1;
        } else {
            if (download_gathered(h, h->ctx[0].vel, vx, vy, 2) != 0) return -1;
        }
    }
    return 0;
}

int gravity_cuda_download(gravity_cuda_t *h, double *x, double *y, double *vx, double *vy)
{
    if (download_enqueue(h, x, y, vx, vy) != 0) return -1;
    if (sync_all(h) != 0) return -1;
    return check_status(h);
}

// ---- the hot path ------------------------------------------------------------------------

// N <= 1024 on one device in STRICT mode: the whole batch in one launch of a multi-step kernel
static int launch_small_strict(gravity_cuda_t *h, double dtd, int64_t nsteps, StepClock *clk)
{
    DeviceCtx &c = h->ctx[0];
    CUDA_TRY(h, cudaSetDevice(c.device));
    int coop = 0;
    CUDA_TRY(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device));
    StepClock dev_clk = *clk;
    dev_clk.ring = c.ring;
    hot_event(h, c);
    if (h->n <= kPairMax) {
        const int threads = (int)(((h->n * h->n + 31) / 32) * 32);
        k_small_strict_steps_pairs<<<1, threads, 0, c.stream>>>(c.pos[h->cur], c.vel, c.mass, (int)h->n, dtd,
                                                                (long long)nsteps, dev_clk);
    } else if (coop) {
        CoopParams P;
        P.pos[0] = c.pos[0];
        P.pos[1] = c.pos[1];
        P.vel = c.vel;
        P.mass = c.mass;
        P.n = (int)h->n;
        P.bodies_per_cta = (int)((h->n + c.num_sms - 1) / c.num_sms);
        P.cur = h->cur;
        P.dt = dtd;
        P.nsteps = (long long)nsteps;
        P.clock = dev_clk;
        const int grid = (int)((h->n + P.bodies_per_cta - 1) / P.bodies_per_cta);
        const size_t smem = coop_smem_bytes(P.n, P.bodies_per_cta);
        CUDA_TRY(h, cudaFuncSetAttribute(k_strict_coop_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        void *args[] = {&P};
        CUDA_TRY(h, cudaLaunchCooperativeKernel((const void *)k_strict_coop_steps, dim3(grid), dim3(kCoopThreads),
                                                args, smem, c.stream));
        h->cur ^= (int)(nsteps & 1);      // the state ends in the other buffer after an odd number of steps
    } else {
        const int threads = (int)(((h->n + 31) / 32) * 32);
        k_small_strict_steps<<<1, threads, 0, c.stream>>>(c.pos[h->cur], c.vel, c.mass, (int)h->n, dtd,
                                                          (long long)nsteps, dev_clk);
    }
    hot_event(h, c);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    for (int64_t k = 0; k < nsteps; ++k) clock_tick(*clk);
    if (clk->capacity > 0) h->path_tail = clk->tail;
    h->step_id += (unsigned long long)nsteps;
    return 0;
}

static int step_impl(gravity_cuda_t *h, int32_t dt, int64_t nsteps, StepClock *clk)
{
    if (!h->uploaded) return fail(h, "step before upload");
    if (nsteps < 0) return fail(h, "nsteps must be >= 0");
    if (h->dead) return fail(h, "the handle is dead: a device-side wait timed out earlier");
    if (nsteps == 0) return 0;
    const double dtd = (double)dt;
    if (effective_kernel(h) == GRAVITY_CUDA_KERNEL_STRICT && h->nranks == 1 && h->n <= kSmallMax)
        return launch_small_strict(h, dtd, nsteps, clk);
    return enqueue_steps(h, dtd, (long long)nsteps, false, clk);
}

int gravity_cuda_step_async(gravity_cuda_t *h, int32_t dt, int64_t nsteps)
{
    if (!h) return fail(nullptr, "step: NULL handle");
    StepClock off;
    memset(&off, 0, sizeof(off));
    off.dt = dt;
    return step_impl(h, dt, nsteps, &off);
}

int gravity_cuda_sync(gravity_cuda_t *h)
{
    if (!h) return fail(nullptr, "sync: NULL handle");
    if (sync_all(h) != 0) return -1;
    return check_status(h);
}

int gravity_cuda_step(gravity_cuda_t *h, int32_t dt, int64_t nsteps)
{
    if (gravity_cuda_step_async(h, dt, nsteps) != 0) return -1;
    return gravity_cuda_sync(h);
}

// ---- PATH ring on the device (sim_gravity.c:1230-1260) -------------------------------------

int gravity_cuda_path_enable(gravity_cuda_t *h, int64_t capacity)
{
    if (!h) return fail(nullptr, "path_enable: NULL handle");
    if (capacity < 0) return fail(h, "path_enable: capacity must be >= 0");
    if (sync_all(h) != 0) return -1;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        cudaFree(c.ring);
        c.ring = nullptr;
        if (capacity > 0) CUDA_TRY(h, cudaMalloc(&c.ring, sizeof(double2) * (size_t)capacity * h->npad));
    }
    h->path_capacity = capacity;
    h->path_tail = 0;
    return 0;
}

int gravity_cuda_step_path(gravity_cuda_t *h, int32_t dt, int64_t nsteps, int64_t sim_time, int32_t *last_day,
                           int64_t *path_tail)
{
    if (!h) return fail(nullptr, "step_path: NULL handle");
    if (!last_day) return fail(h, "step_path: last_day is NULL");
    if (h->path_capacity <= 0) return fail(h, "step_path before path_enable");
    if (dt < 1 || sim_time < 0) return fail(h, "step_path needs dt >= 1 and sim_time >= 0 (the reference never has others)");
    StepClock clk = make_clock(h, h->ctx[0], dt, sim_time, *last_day, true);
    if (step_impl(h, dt, nsteps, &clk) != 0) return -1;
    if (gravity_cuda_sync(h) != 0) return -1;
    *last_day = clk.last_day;
    if (path_tail) *path_tail = h->path_tail;
    return 0;
}

int64_t gravity_cuda_path_samples_in(int32_t dt, int64_t nsteps, int64_t sim_time, int32_t *last_day)
{
    // host-only: the same clock_tick the kernels run, so that a host can size its buffers
    if (dt < 1 || sim_time < 0 || nsteps < 0 || !last_day) return -1;
    StepClock k;
    memset(&k, 0, sizeof(k));
    k.dt = dt;
    k.day = sim_time / 86400;
    k.rem = sim_time % 86400;
    k.capacity = 1;
    k.last_day = *last_day;
    for (int64_t s = 0; s < nsteps; ++s) clock_tick(k);
    *last_day = k.last_day;
    return k.tail;
}

int gravity_cuda_path_read(gravity_cuda_t *h, int64_t first, int64_t count, double *x, double *y)
{
    if (!h) return fail(nullptr, "path_read: NULL handle");
    if (h->path_capacity <= 0) return fail(h, "path_read before path_enable");
    if (count < 0 || first < 0 || first + count > h->path_tail)
        return fail(h, "path_read: samples %lld..%lld were never taken (path_tail = %lld)", (long long)first,
                    (long long)(first + count), (long long)h->path_tail);
    if (h->path_tail - first > h->path_capacity)
        return fail(h, "path_read: sample %lld has been overwritten (ring of %lld rows, path_tail = %lld)",
                    (long long)first, (long long)h->path_capacity, (long long)h->path_tail);
    if (count == 0) return 0;
    if (!x || !y) return fail(h, "path_read: NULL array");
    return no_throw(h, [&]() -> int {
        // every device holds the ring columns of its own shard; rows are copied as they lie
        // (interleaved X,Y) and split on the host -- PATH is display data, a few rows per call
        std::vector<double2> row((size_t)h->chunk);
        if (sync_all(h) != 0) return -1;
        const bool own_only = h->single_process || h->nranks == 1 || h->opt_download_scope == 1;
        for (int64_t k = 0; k < count; ++k) {
            const int64_t r = (first + k) % h->path_capacity;
            if (own_only) {
                for (DeviceCtx &c : h->ctx) {
                    if (c.n_local == 0) continue;
                    CUDA_TRY(h, cudaSetDevice(c.device));
                    CUDA_TRY(h, cudaMemcpy(row.data(), c.ring + (size_t)r * h->npad + c.i0, sizeof(double2) * c.n_local,
                                           cudaMemcpyDeviceToHost));
                    for (int64_t i = 0; i < c.n_local; ++i) {
                        x[k * h->n + c.i0 + i] = row[(size_t)i].x;
                        y[k * h->n + c.i0 + i] = row[(size_t)i].y;
                    }
                }
            } else {
                DeviceCtx &c = h->ctx[0];
                CUDA_TRY(h, cudaSetDevice(c.device));
                double2 *ring_row = c.ring + (size_t)r * h->npad;
                NCCL_TRY(h, ncclAllGather(ring_row + c.i0, ring_row, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
                CUDA_TRY(h, cudaStreamSynchronize(c.stream));
                for (int64_t i0 = 0; i0 < h->n; i0 += h->chunk) {
                    const int64_t cnt = std::min<int64_t>(h->chunk, h->n - i0);
                    CUDA_TRY(h, cudaMemcpy(row.data(), ring_row + i0, sizeof(double2) * cnt, cudaMemcpyDeviceToHost));
                    for (int64_t i = 0; i < cnt; ++i) {
                        x[k * h->n + i0 + i] = row[(size_t)i].x;
                        y[k * h->n + i0 + i] = row[(size_t)i].y;
                    }
                }
            }
        }
        return 0;
    });
}

int gravity_cuda_accel(gravity_cuda_t *h, int32_t dt, double *ax, double *ay)
{
    if (!h) return fail(nullptr, "accel: NULL handle");
    if (!h->uploaded) return fail(h, "accel before upload");
    if (!ax || !ay) return fail(h, "accel: NULL array");
    if (enqueue_steps(h, (double)dt, 1, true, nullptr) != 0) return -1;
    const bool own_only = h->single_process || h->nranks == 1 || h->opt_download_scope == 1;
    if (own_only) {
        if (download_shards(h, src_accel, false, ax, ay, 0) != 0) return -1;
    } else {
        if (download_gathered(h, h->ctx[0].accel, ax, ay, 0) != 0) return -1;
    }
    if (sync_all(h) != 0) return -1;
    return check_status(h);
}

static int step_impl(gravity_cuda_t *h, int32_t dt, int64_t nsteps, StepClock *clk);

int gravity_cuda_step_host(gravity_cuda_t *h, const double *mass, double *x, double *y, double *vx, double *vy,
                           int32_t dt, int64_t nsteps)
{
    if (!h) return fail(nullptr, "step_host: NULL handle");
    // upload, the steps and the download are enqueued back to back and waited for once
    if (upload_enqueue(h, mass, x, y, vx, vy) != 0) return -1;
    StepClock off;
    memset(&off, 0, sizeof(off));
    off.dt = dt;
    if (step_impl(h, dt, nsteps, &off) != 0) return -1;
    if (download_enqueue(h, x, y, vx, vy) != 0) return -1;
    if (sync_all(h) != 0) return -1;
    return check_status(h);
}

int gravity_cuda_step_objects(gravity_cuda_t *h, void **objects, int64_t n, size_t off_x, size_t off_y,
                              size_t off_vx, size_t off_vy, size_t off_mass, int32_t dt, int64_t nsteps)
{
    if (!h) return fail(nullptr, "step_objects: NULL handle");
    if (!objects) return fail(h, "step_objects: NULL objects");
    if (n != h->n) return fail(h, "step_objects: n=%lld but the handle was created for %lld", (long long)n,
                               (long long)h->n);
    return no_throw(h, [&]() -> int {
    // the records are gathered into a page-locked SoA image that the device reads and writes
    // in place: one upload kernel, the steps, one download kernel, one wait per call
    if (!h->aos_stage) {
        CUDA_TRY(h, cudaSetDevice(h->ctx[0].device));
        CUDA_TRY(h, cudaHostAlloc(&h->aos_stage, sizeof(double) * 5 * (size_t)n, cudaHostAllocPortable));
    }
    double *m = h->aos_stage, *x = m + n, *y = x + n, *vx = y + n, *vy = vx + n;
    for (int64_t i = 0; i < n; ++i) {
        const char *o = static_cast<const char *>(objects[i]);
        if (!o) return fail(h, "step_objects: objects[%lld] is NULL", (long long)i);
        memcpy(&m[i], o + off_mass, sizeof(double));
        memcpy(&x[i], o + off_x, sizeof(double));
        memcpy(&y[i], o + off_y, sizeof(double));
        memcpy(&vx[i], o + off_vx, sizeof(double));
        memcpy(&vy[i], o + off_vy, sizeof(double));
    }
    const int scope = h->opt_download_scope;
    h->opt_download_scope = 0;                       // every record is written back, whatever the caller's scope
    const int rc = gravity_cuda_step_host(h, m, x, y, vx, vy, dt, nsteps);
    h->opt_download_scope = scope;
    if (rc != 0) return -1;
    for (int64_t i = 0; i < n; ++i) {
        char *o = static_cast<char *>(objects[i]);
        memcpy(o + off_x, &x[i], sizeof(double));
        memcpy(o + off_y, &y[i], sizeof(double));
        memcpy(o + off_vx, &vx[i], sizeof(double));
        memcpy(o + off_vy, &vy[i], sizeof(double));
    }
    return 0;
    });
}

static int energy_impl(gravity_cuda_t *h, double *kinetic, double *potential)
{
    if (!h) return fail(nullptr, "energy: NULL handle");
    if (!h->uploaded) return fail(h, "energy before upload");
    long double ke = 0, pe = 0;
    std::vector<double> buf;
    for (DeviceCtx &c : h->ctx) {
        if (c.n_local == 0) continue;
        CUDA_TRY(h, cudaSetDevice(c.device));
        EnergyParams P;
        P.pos = c.pos[h->cur];
        P.vel = c.vel;
        P.mass = c.mass;
        P.kin = c.kin;
        P.pot = c.pot;
        P.n = h->n;
        P.i_global0 = c.i0;
        P.n_local = c.n_local;
        const int grid = (int)((c.n_local + kEnergyThreads - 1) / kEnergyThreads);
        if (enqueue_wait_peers(h, c) != 0) return -1;
        k_energy_terms<<<grid, kEnergyThreads, 0, c.stream>>>(P);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
    }
    for (DeviceCtx &c : h->ctx) {
        if (c.n_local == 0) continue;
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaStreamSynchronize(c.stream));
        buf.resize((size_t)2 * c.n_local);
        CUDA_TRY(h, cudaMemcpy(buf.data(), c.kin, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost));
        CUDA_TRY(h, cudaMemcpy(buf.data() + c.n_local, c.pot, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < c.n_local; ++i) {
            ke += buf[i];
            pe += buf[c.n_local + i];
        }
    }
    if (!h->single_process && h->nranks > 1) {
        // combine the per-rank sums (2 doubles) across ranks
        DeviceCtx &c = h->ctx[0];
        double two[2] = {(double)ke, (double)pe};
        CUDA_TRY(h, cudaMemcpy(c.kin, two, sizeof(two), cudaMemcpyHostToDevice));
        NCCL_TRY(h, ncclAllReduce(c.kin, c.kin, 2, ncclDouble, ncclSum, c.comm, c.stream));
        CUDA_TRY(h, cudaStreamSynchronize(c.stream));
        CUDA_TRY(h, cudaMemcpy(two, c.kin, sizeof(two), cudaMemcpyDeviceToHost));
        ke = two[0];
        pe = two[1];
    }
    if (kinetic) *kinetic = (double)ke;
    if (potential) *potential = (double)(-0.5L * (long double)kG * pe);   // each pair was counted twice
    return 0;
}

int gravity_cuda_energy(gravity_cuda_t *h, double *kinetic, double *potential)
{
    return no_throw(h, [&] { return energy_impl(h, kinetic, potential); });
}

int gravity_cuda_timer_start(gravity_cuda_t *h)
{
    if (!h) return fail(nullptr, "timer_start: NULL handle");
    if (sync_all(h) != 0) return -1;
    h->timing = true;
    h->ctx[0].kev_used = 0;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaEventRecord(c.ev_start, c.stream));
    }
    return 0;
}

int gravity_cuda_timer_stop(gravity_cuda_t *h, double *ms)
{
    if (!h) return fail(nullptr, "timer_stop: NULL handle");
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaEventRecord(c.ev_stop, c.stream));
    }
    double worst = 0.0;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaEventSynchronize(c.ev_stop));
        float t = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&t, c.ev_start, c.ev_stop));
        worst = std::max(worst, (double)t);
    }
    h->timing = false;
    if (ms) *ms = worst;
    return 0;
}

int gravity_cuda_kernel_time(gravity_cuda_t *h, double *ms_total, int64_t *launches)
{
    if (!h) return fail(nullptr, "kernel_time: NULL handle");
    DeviceCtx &c = h->ctx[0];
    CUDA_TRY(h, cudaSetDevice(c.device));
    CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    double total = 0.0;
    for (size_t k = 0; k + 1 < c.kev_used; k += 2) {
        float t = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&t, c.kev[k], c.kev[k + 1]));
        total += t;
    }
    if (ms_total) *ms_total = total;
    if (launches) *launches = (int64_t)(c.kev_used / 2);
    return 0;
}

int gravity_cuda_kernel_times(gravity_cuda_t *h, double *ms_out, int64_t cap, int64_t *count)
{
    if (!h) return fail(nullptr, "kernel_times: NULL handle");
    DeviceCtx &c = h->ctx[0];
    CUDA_TRY(h, cudaSetDevice(c.device));
    CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    int64_t k = 0;
    for (size_t e = 0; e + 1 < c.kev_used; e += 2, ++k) {
        if (ms_out && k < cap) {
            float t = 0.f;
            CUDA_TRY(h, cudaEventElapsedTime(&t, c.kev[e], c.kev[e + 1]));
            ms_out[k] = t;
        }
    }
    if (count) *count = k;
    return 0;
}

static int plan_probe_impl(int64_t n, int nranks, int rank, int num_sms, int threads, int bpt, int ctas_per_sm,
                           int32_t *info, int32_t *seg_out, int64_t seg_cap)
{
    if (n < 1 || nranks < 1 || rank < 0 || rank >= nranks || num_sms < 1 || !info)
        return fail(nullptr, "plan_probe: bad argument");
    gravity_cuda geo;
    set_geometry(&geo, n, nranks);
    const int64_t i0 = (int64_t)rank * geo.chunk;
    const int64_t n_local = std::max<int64_t>(0, std::min<int64_t>(geo.chunk, n - i0));
    int t = 0, b = 0, c = 0;
    choose_shape(n_local, (int)((n + kTile - 1) / kTile), num_sms, threads, bpt, ctas_per_sm, t, b, c);
    if (!pick_accum(t, b, c)) return fail(nullptr, "plan_probe: unsupported shape %dx%d", t, b);
    HostPlan plan;
    make_plan(n, n_local, i0, geo.chunk, nranks, num_sms, t, b, c, plan);
    const int32_t vals[10] = {plan.threads, plan.bpt, plan.ctas, plan.n_cta, (int32_t)plan.segs.size(), plan.n_iblock,
                              plan.n_tile, plan.local_tile_begin, plan.local_tile_end, (int32_t)n_local};
    memcpy(info, vals, sizeof(vals));
    if (seg_out) {
        for (int k = 0; k < plan.n_cta; ++k) {
            for (int sidx = plan.cta_begin[k]; sidx < plan.cta_begin[k + 1]; ++sidx) {
                if (sidx >= seg_cap) break;
                const Segment &sg = plan.segs[sidx];
                int32_t *o = seg_out + (size_t)sidx * 5;
                o[0] = k; o[1] = sg.iblock; o[2] = sg.tile_begin; o[3] = sg.tile_end; o[4] = sg.slot;
            }
        }
    }
    return 0;
}

int gravity_cuda_plan_probe(int64_t n, int nranks, int rank, int num_sms, int threads, int bpt, int ctas_per_sm,
                            int32_t *info, int32_t *seg_out, int64_t seg_cap)
{
    return no_throw(nullptr, [&] {
        return plan_probe_impl(n, nranks, rank, num_sms, threads, bpt, ctas_per_sm, info, seg_out, seg_cap);
    });
}

int gravity_cuda_trace_read(gravity_cuda_t *h, uint64_t *stamps, int64_t cap_ctas, int64_t *n_cta)
{
    if (!h || !n_cta) return fail(h, "trace_read: NULL argument");
    DeviceCtx &c = h->ctx[0];
    if (!c.trace) return fail(h, "trace_read: set option \"trace\" = 1 before the first step");
    CUDA_TRY(h, cudaSetDevice(c.device));
    CUDA_TRY(h, cudaStreamSynchronize(c.stream));
    *n_cta = c.phase[0].n_cta;
    const int64_t rows = std::min<int64_t>(cap_ctas, c.phase[0].n_cta);
    if (stamps && rows > 0)
        CUDA_TRY(h, cudaMemcpy(stamps, c.trace, sizeof(uint64_t) * (size_t)rows * kTraceEvents, cudaMemcpyDeviceToHost));
    return 0;
}

int gravity_cuda_launch_count(const gravity_cuda_t *h, int64_t *launches)
{
    if (!h || !launches) return -1;
    *launches = h->launches;
    return 0;
}

int gravity_cuda_fp64_peak(int device, double *tflops, double *sm_mhz)
{
    if (check_device(nullptr, device) != 0) return -1;
    CUDA_TRY(nullptr, cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(nullptr, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double *sink = nullptr;
    long long *cycles = nullptr;
    CUDA_TRY(nullptr, cudaMalloc(&sink, sizeof(double)));
    CUDA_TRY(nullptr, cudaMalloc(&cycles, sizeof(long long)));
    cudaEvent_t e0, e1;
    CUDA_TRY(nullptr, cudaEventCreate(&e0));
    CUDA_TRY(nullptr, cudaEventCreate(&e1));
    const int grid = sms * 4, threads = 256;
    int iters = 2000;
    double best = 0.0, best_mhz = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(nullptr, cudaMemset(cycles, 0, sizeof(long long)));
        CUDA_TRY(nullptr, cudaEventRecord(e0));
        k_dfma_peak<<<grid, threads>>>(sink, iters, 0.25, 1e-12, cycles);
        CUDA_TRY(nullptr, cudaEventRecord(e1));
        CUDA_TRY(nullptr, cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(nullptr, cudaEventElapsedTime(&ms, e0, e1));
        long long cyc = 0;
        CUDA_TRY(nullptr, cudaMemcpy(&cyc, cycles, sizeof(cyc), cudaMemcpyDeviceToHost));
        const double flops = 2.0 * (double)grid * threads * (double)iters * 8.0 * kPeakChains;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep == 0 && ms < 20.f) { iters = (int)(iters * 40.f / std::max(ms, 0.01f)); continue; }
        if (tf > best) { best = tf; best_mhz = (double)cyc / (ms * 1e-3) / 1e6; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cycles);
    if (tflops) *tflops = best;
    if (sm_mhz) *sm_mhz = best_mhz;
    return 0;
}

}  // extern "C"
