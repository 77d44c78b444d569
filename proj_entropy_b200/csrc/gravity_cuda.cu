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
    unsigned int       *status = nullptr;
    double2            *peer_pos[2][kMaxPeers] = {};
    unsigned long long *peer_flags[kMaxPeers] = {};
    bool                ipc_opened[kMaxPeers] = {};
    double  *stage = nullptr;      // 4 x npad doubles: SoA staging for upload / download
    double  *kin = nullptr, *pot = nullptr;
    double2 *partial = nullptr;
    Segment *seg = nullptr;
    int32_t *cta_seg_begin = nullptr;
    int32_t *iblock_seg_begin = nullptr;

    int64_t i0 = 0;          // global index of local body 0
    int64_t n_local = 0;     // real bodies in this shard

    // launch plan of the tiled kernel
    struct Phase { int n_cta = 0; int cta_seg_offset = 0; };
    Phase phase[1];
    int n_iblock = 0;
    int local_tile_begin = 0, local_tile_end = 0;   // j-tiles of this rank's own shard
    int32_t *iblock_arrivals = nullptr;             // [n_iblock], see k_tiled_accumulate
    int32_t *iblocks_done = nullptr;
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
    bool p2p_connected = false;
    unsigned long long step_id = 0;  // commits published so far (peer-memory exchange)
    int64_t launches = 0;
    bool timing = false;
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

typedef void (*accum_kernel_t)(const AccumParams);

template <int T, int R>
accum_kernel_t accum_for_minb(int minb)
{
    switch (minb) {
    case 1: return k_tiled_accumulate<T, R, 1>;
    case 2: return k_tiled_accumulate<T, R, 2>;
    case 3: return k_tiled_accumulate<T, R, 3>;
    default: return k_tiled_accumulate<T, R, 4>;
    }
}

accum_kernel_t pick_accum(int threads, int bpt, int minb)
{
    if (threads == 128) {
        if (bpt == 1) return accum_for_minb<128, 1>(minb);
        if (bpt == 2) return accum_for_minb<128, 2>(minb);
        if (bpt == 4) return accum_for_minb<128, 4>(minb);
    } else if (threads == 256) {
        if (bpt == 1) return accum_for_minb<256, 1>(minb);
        if (bpt == 2) return accum_for_minb<256, 2>(minb);
        if (bpt == 4) return accum_for_minb<256, 4>(minb);
    }
    return nullptr;
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
            if (score > best) {
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
        accum_kernel_t probe = pick_accum(threads, bpt, ctas);
        if (!probe) return fail(h, "unsupported tiled kernel shape threads=%d bodies_per_thread=%d", threads, bpt);
        int occ = 0;
        CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, threads, 0));
        if (occ < 1) return fail(h, "tiled kernel does not fit on an SM");
        HostPlan plan;
        make_plan(h->n, c.n_local, c.i0, h->chunk, h->nranks, c.num_sms, threads, bpt, std::min(occ, ctas), plan);
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

        cudaFree(c.seg); cudaFree(c.cta_seg_begin); cudaFree(c.iblock_seg_begin); cudaFree(c.partial);
        cudaFree(c.iblock_arrivals); cudaFree(c.iblocks_done);
        c.seg = nullptr; c.cta_seg_begin = nullptr; c.iblock_seg_begin = nullptr; c.partial = nullptr;
        c.iblock_arrivals = nullptr; c.iblocks_done = nullptr;
        CUDA_TRY(h, cudaMalloc(&c.iblock_arrivals, sizeof(int32_t) * std::max(1, c.n_iblock)));
        CUDA_TRY(h, cudaMalloc(&c.iblocks_done, sizeof(int32_t)));
        CUDA_TRY(h, cudaMemset(c.iblock_arrivals, 0, sizeof(int32_t) * std::max(1, c.n_iblock)));
        CUDA_TRY(h, cudaMemset(c.iblocks_done, 0, sizeof(int32_t)));
        CUDA_TRY(h, cudaMalloc(&c.seg, sizeof(Segment) * std::max<size_t>(1, plan.segs.size())));
        CUDA_TRY(h, cudaMalloc(&c.cta_seg_begin, sizeof(int32_t) * std::max<size_t>(1, plan.cta_begin.size())));
        CUDA_TRY(h, cudaMalloc(&c.iblock_seg_begin, sizeof(int32_t) * plan.ib_begin.size()));
        CUDA_TRY(h, cudaMalloc(&c.partial, sizeof(double2) * std::max<size_t>(1, plan.segs.size()) * IB));
        CUDA_TRY(h, cudaMemcpy(c.seg, plan.segs.data(), sizeof(Segment) * plan.segs.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(c.cta_seg_begin, plan.cta_begin.data(), sizeof(int32_t) * plan.cta_begin.size(),
                               cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(c.iblock_seg_begin, plan.ib_begin.data(), sizeof(int32_t) * plan.ib_begin.size(),
                               cudaMemcpyHostToDevice));
    }
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

// ---- one step on every local device -----------------------------------------------

CommitParams commit_params(gravity_cuda_t *h, DeviceCtx &c, double dt, bool export_accel)
{
    CommitParams P;
    memset(&P, 0, sizeof(P));
    P.pos_cur = c.pos[h->cur];
    P.pos_next = c.pos[h->cur ^ 1];
    P.vel = c.vel;
    P.mass = c.mass;
    P.partial = c.partial;
    P.iblock_seg_begin = c.iblock_seg_begin;
    P.iblock_arrivals = c.iblock_arrivals;
    P.iblocks_done = c.iblocks_done;
    P.n_iblock = c.n_iblock;
    P.accel_out = c.accel;
    P.n = h->n;
    P.i_global0 = c.i0;
    P.n_local = c.n_local;
    P.iblock_size = c.threads * c.bpt;
    P.export_accel = export_accel ? 1 : 0;
    P.dt = dt;
    const bool p2p = h->opt_exchange == 1 && h->nranks > 1;
    P.wait_flags = p2p ? c.flags : nullptr;
    P.wait_step = h->step_id;
    P.px.enabled = (p2p && !export_accel) ? 1 : 0;
    P.px.rank = c.rank;
    P.px.nranks = h->nranks;
    P.px.flags = c.flags;
    P.px.commit_counter = c.commit_counter;
    P.px.status = c.status;
    P.px.step = h->step_id + 1;
    for (int r = 0; r < h->nranks && r < kMaxPeers; ++r) {
        P.px.peer_next[r] = c.peer_pos[h->cur ^ 1][r];
        P.px.peer_flags[r] = c.peer_flags[r];
    }
    return P;
}

// The one launch of a tiled step: pairs, reduction, repair, kick + drift, hand-over.
int launch_tiled(gravity_cuda_t *h, DeviceCtx &c, double dt, bool export_accel)
{
    const DeviceCtx::Phase &ph = c.phase[0];
    if (ph.n_cta == 0) return 0;
    AccumParams P;
    memset(&P, 0, sizeof(P));
    P.pos = c.pos[h->cur];
    P.mass = c.mass;
    P.vel = c.vel;
    P.partial = c.partial;
    P.seg = c.seg;
    P.cta_seg_begin = c.cta_seg_begin + ph.cta_seg_offset;
    P.tile_mass = h->opt_mass_fold ? c.tile_mass : nullptr;
    P.i_global0 = c.i0;
    P.n_local = c.n_local;
    P.dt = dt;
    const bool p2p = h->opt_exchange == 1 && h->nranks > 1;
    P.wait_flags = p2p ? c.flags : nullptr;
    P.wait_status = c.status;
    P.wait_step = h->step_id;
    P.wait_rank = c.rank;
    P.wait_nranks = h->nranks;
    P.local_tile_begin = c.local_tile_begin;
    P.local_tile_end = c.local_tile_end;
    P.cm = commit_params(h, c, dt, export_accel);
    accum_kernel_t k = pick_accum(c.threads, c.bpt, c.minb);
    hot_event(h, c);
    k<<<ph.n_cta, c.threads, 0, c.stream>>>(P);
    hot_event(h, c);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

// Evaluate sim_gravity.c:1183-1222 once on all local devices, either advancing
// the state (export_accel == false) or writing AX,AY into ctx.accel.
int enqueue_evaluation(gravity_cuda_t *h, double dt, bool export_accel, bool gather_after)
{
    const int kernel = effective_kernel(h);
    if (kernel == GRAVITY_CUDA_KERNEL_TILED && !h->planned) {
        if (build_plan(h) != 0) return -1;
    }
    const bool p2p = h->opt_exchange == 1 && h->nranks > 1;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        if (c.n_local == 0) {
            // an empty shard still has to tell its peers that it has nothing to send
            if (p2p && !export_accel) {
                const CommitParams P = commit_params(h, c, dt, false);
                k_publish_only<<<1, 32, 0, c.stream>>>(P.px);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
            continue;
        }
        if (kernel == GRAVITY_CUDA_KERNEL_TILED) {
            if (launch_tiled(h, c, dt, export_accel) != 0) return -1;
        } else {
            const CommitParams P = commit_params(h, c, dt, export_accel);
            const int grid = (int)((c.n_local + kStrictThreads - 1) / kStrictThreads);
            hot_event(h, c);
            k_strict_rows<<<grid, kStrictThreads, 0, c.stream>>>(P);
            hot_event(h, c);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
    }
    if (!export_accel && gather_after && h->nranks > 1) {
        if (p2p) {
            h->step_id++;      // the kernels stored into the peers and publish this step
        } else {
            // barrier B2 + the NEXT_* -> state copy (sim_gravity.c:1226, :1242-1245)
            // become one in-place all-gather of the new positions.
            NCCL_TRY(h, ncclGroupStart());
            for (DeviceCtx &c : h->ctx) {
                double2 *next = c.pos[h->cur ^ 1];
                NCCL_TRY(h, ncclAllGather(next + c.i0, next, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
            }
            NCCL_TRY(h, ncclGroupEnd());
        }
    }
    if (!export_accel) h->cur ^= 1;
    return 0;
}

// peer-memory exchange: make the stream wait until every peer's positions of
// the current step have landed in this device's pos[cur]
int enqueue_wait_peers(gravity_cuda_t *h, DeviceCtx &c)
{
    if (!(h->opt_exchange == 1 && h->nranks > 1)) return 0;
    k_wait_peers<<<1, 32, 0, c.stream>>>(c.flags, h->step_id, c.rank, h->nranks, c.status);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

int check_exchange_status(gravity_cuda_t *h)
{
    if (!(h->opt_exchange == 1 && h->nranks > 1)) return 0;
    for (DeviceCtx &c : h->ctx) {
        unsigned int st = 0;
        CUDA_TRY(h, cudaSetDevice(c.device));
        CUDA_TRY(h, cudaMemcpy(&st, c.status, sizeof(st), cudaMemcpyDeviceToHost));
        if (st != 0) return fail(h, "peer-memory exchange timed out on rank %d: a peer never published its step", c.rank);
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
    c.i0 = (int64_t)c.rank * h->chunk;
    c.n_local = std::max<int64_t>(0, std::min<int64_t>(h->chunk, h->n - c.i0));
    CUDA_TRY(h, cudaMalloc(&c.pos[0], sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.pos[1], sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.mass, sizeof(double) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.vel, sizeof(double2) * h->chunk));
    CUDA_TRY(h, cudaMalloc(&c.accel, sizeof(double2) * h->chunk));
    CUDA_TRY(h, cudaMalloc(&c.stage, sizeof(double) * 4 * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.tile_mass, sizeof(double) * (h->npad / kTile)));
    CUDA_TRY(h, cudaMalloc(&c.gather_tmp, sizeof(double2) * h->npad));
    CUDA_TRY(h, cudaMalloc(&c.flags, sizeof(unsigned long long) * kMaxPeers));
    CUDA_TRY(h, cudaMalloc(&c.commit_counter, sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&c.status, sizeof(unsigned int)));
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
    if (nranks < 1 || nranks > 64) return fail(nullptr, "number of GPUs must be 1..64 (got %d)", nranks);
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
    for (DeviceCtx &c : h->ctx) {
        cudaSetDevice(c.device);
        if (c.stream) cudaStreamSynchronize(c.stream);
        if (c.comm) ncclCommDestroy(c.comm);
        cudaFree(c.pos[0]); cudaFree(c.pos[1]); cudaFree(c.mass); cudaFree(c.vel); cudaFree(c.accel);
        cudaFree(c.kin); cudaFree(c.pot); cudaFree(c.stage); cudaFree(c.tile_mass);
        for (int r = 0; r < kMaxPeers; ++r) {
            if (c.ipc_opened[r]) {
                cudaIpcCloseMemHandle(c.peer_pos[0][r]);
                cudaIpcCloseMemHandle(c.peer_pos[1][r]);
                cudaIpcCloseMemHandle(c.peer_flags[r]);
            }
        }
        cudaFree(c.gather_tmp); cudaFree(c.flags); cudaFree(c.commit_counter); cudaFree(c.status); cudaFree(c.partial); cudaFree(c.seg);
        cudaFree(c.cta_seg_begin); cudaFree(c.iblock_seg_begin);
        cudaFree(c.iblock_arrivals); cudaFree(c.iblocks_done);
        for (cudaEvent_t e : c.kev) cudaEventDestroy(e);
        if (c.ev_start) cudaEventDestroy(c.ev_start);
        if (c.ev_stop) cudaEventDestroy(c.ev_stop);
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
        h->opt_exchange = (int)value;
    } else {
        return fail(h, "unknown option '%s'", key);
    }
    h->planned = false;
    return 0;
}

int gravity_cuda_upload(gravity_cuda_t *h, const double *mass, const double *x, const double *y,
                        const double *vx, const double *vy)
{
    if (!h) return fail(nullptr, "upload: NULL handle");
    if (!mass || !x || !y || !vx || !vy) return fail(h, "upload: NULL array");
    if (rank_barrier(h) != 0) return -1;
    const int64_t n = h->n, np = h->npad;
    h->cur = 0;
    for (DeviceCtx &c : h->ctx) {
        CUDA_TRY(h, cudaSetDevice(c.device));
        double *sx = c.stage, *sy = c.stage + np, *svx = c.stage + 2 * np, *svy = c.stage + 3 * np;
        CUDA_TRY(h, cudaMemcpyAsync(sx, x, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
        CUDA_TRY(h, cudaMemcpyAsync(sy, y, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
        CUDA_TRY(h, cudaMemcpyAsync(c.mass, mass, sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
        k_pack_pairs<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(sx, sy, c.pos[0], n);
        const int n_tile = (int)(np / kTile);
        k_tile_mass<<<(n_tile + 3) / 4, 128, 0, c.stream>>>(c.mass, c.tile_mass, n_tile);
        h->launches += 2;
        if (c.n_local > 0) {
            CUDA_TRY(h, cudaMemcpyAsync(svx, vx + c.i0, sizeof(double) * c.n_local, cudaMemcpyHostToDevice, c.stream));
            CUDA_TRY(h, cudaMemcpyAsync(svy, vy + c.i0, sizeof(double) * c.n_local, cudaMemcpyHostToDevice, c.stream));
            k_pack_pairs<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(svx, svy, c.vel, c.n_local);
            h->launches++;
        }
        CUDA_TRY(h, cudaGetLastError());
        }
    if (sync_all(h) != 0) return -1;
    h->uploaded = true;
    return 0;
}

int gravity_cuda_download(gravity_cuda_t *h, double *x, double *y, double *vx, double *vy)
{
    if (!h) return fail(nullptr, "download: NULL handle");
    if (!h->uploaded) return fail(h, "download before upload");
    if (sync_all(h) != 0) return -1;
    const int64_t n = h->n, np = h->npad;
    if (x || y) {
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
    if (vx || vy) {
        if (h->single_process || h->nranks == 1) {
            for (DeviceCtx &c : h->ctx) {
                if (c.n_local == 0) continue;
                CUDA_TRY(h, cudaSetDevice(c.device));
                double *svx = c.stage + 2 * np, *svy = c.stage + 3 * np;
                k_unpack_pairs<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(c.vel, svx, svy, c.n_local);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
                if (vx) CUDA_TRY(h, cudaMemcpyAsync(vx + c.i0, svx, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
                if (vy) CUDA_TRY(h, cudaMemcpyAsync(vy + c.i0, svy, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
            }
        } else {
            // rank mode: gather the velocity shards through the spare position buffer
            DeviceCtx &c = h->ctx[0];
            CUDA_TRY(h, cudaSetDevice(c.device));
            double2 *tmp = c.gather_tmp;
            double *svx = c.stage + 2 * np, *svy = c.stage + 3 * np;
            CUDA_TRY(h, cudaMemcpyAsync(tmp + c.i0, c.vel, sizeof(double2) * h->chunk, cudaMemcpyDeviceToDevice, c.stream));
            NCCL_TRY(h, ncclAllGather(tmp + c.i0, tmp, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
            k_unpack_pairs<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(tmp, svx, svy, n);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            if (vx) CUDA_TRY(h, cudaMemcpyAsync(vx, svx, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
            if (vy) CUDA_TRY(h, cudaMemcpyAsync(vy, svy, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
        }
    }
    if (sync_all(h) != 0) return -1;
    return check_exchange_status(h);
}

int gravity_cuda_step_async(gravity_cuda_t *h, int32_t dt, int64_t nsteps)
{
    if (!h) return fail(nullptr, "step: NULL handle");
    if (!h->uploaded) return fail(h, "step before upload");
    if (nsteps < 0) return fail(h, "nsteps must be >= 0");
    if (nsteps == 0) return 0;
    const double dtd = (double)dt;
    const int kernel = effective_kernel(h);
    if (kernel == GRAVITY_CUDA_KERNEL_STRICT && h->nranks == 1 && h->n <= kSmallMax) {
        // the whole batch in one single-CTA launch
        DeviceCtx &c = h->ctx[0];
        CUDA_TRY(h, cudaSetDevice(c.device));
        int coop = 0;
        CUDA_TRY(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device));
        hot_event(h, c);
        if (h->n <= kPairMax) {
            const int threads = (int)(((h->n * h->n + 31) / 32) * 32);
            k_small_strict_steps_pairs<<<1, threads, 0, c.stream>>>(c.pos[h->cur], c.vel, c.mass, (int)h->n, dtd,
                                                                    (long long)nsteps);
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
                                                              (long long)nsteps);
        }
        hot_event(h, c);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        return 0;
    }
    for (int64_t s = 0; s < nsteps; ++s) {
        if (enqueue_evaluation(h, dtd, false, true) != 0) return -1;
    }
    return 0;
}

int gravity_cuda_sync(gravity_cuda_t *h)
{
    if (!h) return fail(nullptr, "sync: NULL handle");
    if (sync_all(h) != 0) return -1;
    return check_exchange_status(h);
}

int gravity_cuda_step(gravity_cuda_t *h, int32_t dt, int64_t nsteps)
{
    if (gravity_cuda_step_async(h, dt, nsteps) != 0) return -1;
    return gravity_cuda_sync(h);
}

int gravity_cuda_accel(gravity_cuda_t *h, int32_t dt, double *ax, double *ay)
{
    if (!h) return fail(nullptr, "accel: NULL handle");
    if (!h->uploaded) return fail(h, "accel before upload");
    if (!ax || !ay) return fail(h, "accel: NULL array");
    if (enqueue_evaluation(h, (double)dt, true, false) != 0) return -1;
    if (sync_all(h) != 0) return -1;
    const int64_t n = h->n, np = h->npad;
    if (h->single_process || h->nranks == 1) {
        for (DeviceCtx &c : h->ctx) {
            if (c.n_local == 0) continue;
            CUDA_TRY(h, cudaSetDevice(c.device));
            double *sx = c.stage, *sy = c.stage + np;
            k_unpack_pairs<<<(unsigned)((c.n_local + 255) / 256), 256, 0, c.stream>>>(c.accel, sx, sy, c.n_local);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            CUDA_TRY(h, cudaMemcpyAsync(ax + c.i0, sx, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
            CUDA_TRY(h, cudaMemcpyAsync(ay + c.i0, sy, sizeof(double) * c.n_local, cudaMemcpyDeviceToHost, c.stream));
        }
    } else {
        DeviceCtx &c = h->ctx[0];
        CUDA_TRY(h, cudaSetDevice(c.device));
        double2 *tmp = c.gather_tmp;
        double *sx = c.stage, *sy = c.stage + np;
        CUDA_TRY(h, cudaMemcpyAsync(tmp + c.i0, c.accel, sizeof(double2) * h->chunk, cudaMemcpyDeviceToDevice, c.stream));
        NCCL_TRY(h, ncclAllGather(tmp + c.i0, tmp, (size_t)h->chunk * 2, ncclDouble, c.comm, c.stream));
        k_unpack_pairs<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(tmp, sx, sy, n);
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaMemcpyAsync(ax, sx, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
        CUDA_TRY(h, cudaMemcpyAsync(ay, sy, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    }
    return sync_all(h);
}

int gravity_cuda_step_objects(gravity_cuda_t *h, void **objects, int64_t n, size_t off_x, size_t off_y,
                              size_t off_vx, size_t off_vy, size_t off_mass, int32_t dt, int64_t nsteps)
{
    if (!h) return fail(nullptr, "step_objects: NULL handle");
    if (!objects) return fail(h, "step_objects: NULL objects");
    if (n != h->n) return fail(h, "step_objects: n=%lld but the handle was created for %lld", (long long)n,
                               (long long)h->n);
    return no_throw(h, [&]() -> int {
    std::vector<double> soa((size_t)5 * n);
    double *m = soa.data(), *x = m + n, *y = x + n, *vx = y + n, *vy = vx + n;
    for (int64_t i = 0; i < n; ++i) {
        const char *o = static_cast<const char *>(objects[i]);
        if (!o) return fail(h, "step_objects: objects[%lld] is NULL", (long long)i);
        memcpy(&m[i], o + off_mass, sizeof(double));
        memcpy(&x[i], o + off_x, sizeof(double));
        memcpy(&y[i], o + off_y, sizeof(double));
        memcpy(&vx[i], o + off_vx, sizeof(double));
        memcpy(&vy[i], o + off_vy, sizeof(double));
    }
    if (gravity_cuda_upload(h, m, x, y, vx, vy) != 0) return -1;
    if (gravity_cuda_step(h, dt, nsteps) != 0) return -1;
    if (gravity_cuda_download(h, x, y, vx, vy) != 0) return -1;
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
