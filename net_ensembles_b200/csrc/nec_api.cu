// C ABI of net_ensembles_cuda (include/net_ensembles_cuda.h): context, device CSR, the
// vertex_load entry points and their workspace management.  All compute happens in the
// kernels of nec_kernels.cuh / nec_small.cuh; there is no host implementation of the
// algorithm here and no fallback when CUDA is unavailable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "../../include/net_ensembles_cuda.h"
#include "nec_kernels.cuh"
#include "nec_small.cuh"
#include "nec_chain.cuh"
#include "nec_sample.cuh"
#include "nec_mid.cuh"
#include "nec_wide.cuh"

namespace {

thread_local std::string g_init_error;

struct Arena {                       // per-slot scratch of the batched engine
    void *base = nullptr;
    size_t bytes = 0;
    // layout parameters the arena was built for
    uint32_t n = 0, n_slots = 0, dist_bytes = 0, max_levels = 0, ksrc = 0;
    size_t qcap = 0;
    // carved pointers
    uint8_t *dist = nullptr; size_t dist_stride = 0;
    double *q = nullptr;     size_t q_stride = 0;
    uint8_t *mask = nullptr; size_t mask_stride = 0;    // bytes per slot
    double *bslot = nullptr; size_t bslot_stride = 0;
    uint32_t *queue_v = nullptr; uint8_t *queue_m = nullptr; size_t queue_stride = 0;   // entries per slot
    uint32_t *lvl = nullptr; size_t lvl_stride = 0;
};

struct WideArena {                   // per-slot scratch of the compact (wide) engine, nec_wide.cuh
    void *base = nullptr;
    size_t bytes = 0;
    uint32_t n = 0, n_slots = 0, kw = 0, max_levels = 0;
    size_t qcap = 0;
    uint8_t *rec = nullptr;  size_t rec_stride = 0;      // bytes per slot
    uint8_t *lr = nullptr;   size_t lr_stride = 0;       // bytes per slot
    double *qc = nullptr;    size_t qc_stride = 0;       // doubles per slot
    double *bslot = nullptr; size_t bslot_stride = 0;
    uint32_t *queue_v = nullptr, *queue_p = nullptr; uint8_t *queue_m = nullptr; size_t queue_stride = 0;
    uint32_t *lvl = nullptr; size_t lvl_stride = 0;
    uint32_t *bq = nullptr, *bcnt = nullptr; size_t bq_stride = 0;   // 256 sources per batch: the backward work list and its per-level counts
};

}  // namespace

struct nec_graph {
    uint32_t n = 0;
    uint64_t nnz = 0;
    uint32_t *d_row_ptr = nullptr;   // n+1
    uint32_t *d_col_idx = nullptr;   // nnz
    uint32_t max_degree = 0;
    uint4 *d_slab = nullptr;         // n rows of slab_w u32 (mid engine), nullptr if n > kMidMaxN
    uint8_t *d_deg8 = nullptr;       // min(degree, 255) per vertex (mid engine)
    int slab_w = 0;
    std::vector<nec_graph *> replicas;   // multi-device context: the same CSR on devices 1..
    // Which engine runs large calls on a graph that fits the mid engine (engine_hint below): 0 not probed yet, 2 mid, 4 wide;
    // and the probe's finding, queue entries per vertex of one 256-source batch
    mutable int engine_hint = 0;
    mutable float probe_entries = 0.f;
};

struct nec_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    std::string err;
    uint64_t ws_limit = 0;
    Arena arena;
    WideArena wide;
    int wide_cap[3][2][9];                            // resident clusters per (width 64/128/256, CTA 512/1024 threads, team size); -1 = not asked yet
    // small device buffers reused across calls
    uint32_t *d_sources = nullptr; size_t d_sources_cap = 0;
    uint8_t *d_status = nullptr;   size_t d_status_cap = 0;
    uint32_t *d_worklist = nullptr; size_t d_worklist_cap = 0;
    unsigned long long *d_counters = nullptr;       // 8
    double *d_out = nullptr; size_t d_out_cap = 0;  // result staging for host-pointer calls
    double *d_keep = nullptr; size_t d_keep_cap = 0;  // pass-1 sums while the 32-bit distance pass runs
    double *d_tmp = nullptr; size_t d_tmp_cap = 0;    // partial load of the sources the mid engine handed back
    // nec_distance_measures outputs (per source) and the hand-back subset's, kept across calls
    uint32_t *d_mecc = nullptr, *d_mreach = nullptr, *d_pecc = nullptr, *d_preach = nullptr, *d_scatter = nullptr;
    unsigned long long *d_msum = nullptr, *d_psum = nullptr;
    size_t d_mecc_cap = 0, d_mreach_cap = 0, d_msum_cap = 0, d_pecc_cap = 0, d_preach_cap = 0, d_psum_cap = 0, d_scatter_cap = 0;
    int occ_u8 = 0, occ_u32 = 0;
    void *mid_base = nullptr; size_t mid_bytes = 0;   // arena of the one-source-per-CTA engine
    uint8_t *d_src_status = nullptr; size_t d_src_status_cap = 0;   // mid engine: [8 counters | per-source status], one memset and one copy back per call
    bool mid_attr_set = false;
    uint8_t *h_pin = nullptr; size_t h_pin_cap = 0;      // pinned host staging of a call's small transfers (source list in, verdicts + counters out)
    struct { uint32_t n = 0; int slab_w = 0, threads = 0, occ = 0; } mid_occ;   // occupancy of the last mid-engine shape asked for
    int force_engine = 0;                             // 0 auto, 1 batched only, 2 mid whenever it fits
    bool probe_mode = false;                          // engine_hint: a probe batch is running (its hand-backs are not re-run)
    double *dbg_sigma = nullptr;                      // nec_bfs_debug: plane for the path counts of the debug instantiation
    uint32_t pull_mode = 1;                           // batched forward pass: 0 push only, 1 per-level cost model, 2 pull only (NEC_PULL, tests)
    bool wide_ctas = false;                           // batched engine: 1024-thread CTAs, one per SM (set per call)
    uint32_t team_size = 1;                           // batched engine: CTAs per slot (thread-block cluster), set per call
    int team_cap[2][9];                               // resident clusters per (batch width 32/64, team size); -1 = not asked yet
    nec::SmallWorkspace small;
    void *h_stage = nullptr; size_t h_stage_cap = 0;  // pinned staging of an upload's 32-bit CSR
    uint8_t *d_sample = nullptr; size_t d_sample_cap = 0;   // scratch of nec_erc_randomize
    // multi-device context (nec_init with n_devices > 1): this context works on devices[0]; `peers` are full
    // single-device contexts for devices[1..], `comms[i]` the NCCL communicator of device i (0 = this one)
    std::vector<nec_ctx *> peers;
    std::vector<ncclComm_t> comms;
    nec_stats stats{};
};

namespace {

int fail(nec_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_init_error = buf;
    return code;
}

#define NEC_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? NEC_ENOMEM : NEC_ECUDA,          \
                        "%s failed: %s", #call, cudaGetErrorString(e__));                        \
    } while (0)


// No C++ exception may cross the C ABI: host allocations inside an entry point report NEC_ENOMEM instead.
#define NEC_CATCH_ALL(ctx)                                                                        \
    catch (const std::bad_alloc &) { return fail(ctx, NEC_ENOMEM, "out of host memory"); }        \
    catch (const std::exception &ex) { return fail(ctx, NEC_EINTERNAL, "unexpected exception: %s", ex.what()); }

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <class T>
int grow(nec_ctx *ctx, T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return NEC_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    NEC_CUDA(ctx, cudaMalloc((void **)&ptr, need * sizeof(T)));
    cap = need;
    return NEC_OK;
}

int grow_pinned(nec_ctx *ctx, size_t need) {
    if (need <= ctx->h_pin_cap) return NEC_OK;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    const size_t cap = std::max<size_t>(need + need / 2, 1 << 16);
    NEC_CUDA(ctx, cudaMallocHost((void **)&ctx->h_pin, cap));
    ctx->h_pin_cap = cap;
    return NEC_OK;
}

size_t per_slot_bytes(uint32_t n, uint32_t dist_bytes, size_t qcap, uint32_t max_levels, uint32_t ksrc) {
    const size_t mask_b = ksrc / 8;
    const size_t qpitch = align_up(qcap, 64);
    size_t b = 0;
    b += align_up((size_t)n * ksrc * dist_bytes, 256);
    b += align_up((size_t)n * ksrc * sizeof(double), 256);
    b += align_up((size_t)n * 4 * mask_b, 256);
    b += align_up((size_t)n * sizeof(double), 256);
    b += qpitch * 4 + qpitch * mask_b;
    b += align_up(((size_t)max_levels + 2) * sizeof(uint32_t), 256);
    return b;
}

// Slots and queue capacity the workspace budget allows for (n, distance width, batch width, wanted slots).
struct ArenaPlan { uint32_t slots; size_t qcap, per_slot, budget; };
int plan_arena(nec_ctx *ctx, uint32_t n, uint32_t dist_bytes, uint32_t want_slots, uint32_t ksrc, ArenaPlan &out, bool full_queue = false) {
    const uint32_t max_levels = dist_bytes == 1 ? 254u : n + 1;
    size_t free_b = 0, total_b = 0;
    NEC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = ctx->ws_limit ? (size_t)ctx->ws_limit : (size_t)((double)(free_b + ctx->arena.bytes + ctx->wide.bytes) * 0.8);
    // queue capacity: every vertex can sit on at most 32 (64) distinct levels of one batch
    size_t qcap = (size_t)n * ksrc + ksrc;
    uint32_t slots = want_slots;
    size_t ps = per_slot_bytes(n, dist_bytes, qcap, max_levels, ksrc);
    if (ps * slots > budget) {
        // prefer keeping one slot per SM; shrink the queue bound (overflow is detected) first
        const uint32_t floor_slots = std::min<uint32_t>(slots, (uint32_t)ctx->prop.multiProcessorCount);
        while (!full_queue && ps * floor_slots > budget && qcap > (size_t)n * 8 + ksrc) {
            qcap = std::max((size_t)n * 8 + ksrc, qcap / 2);
            ps = per_slot_bytes(n, dist_bytes, qcap, max_levels, ksrc);
        }
        slots = (uint32_t)std::min<size_t>(slots, budget / ps);
    }
    out = ArenaPlan{slots, qcap, ps, budget};
    return NEC_OK;
}

// Build (or reuse) the arena for (n, distance width, wanted slots).
int ensure_arena(nec_ctx *ctx, uint32_t n, uint32_t dist_bytes, uint32_t want_slots, uint32_t ksrc = nec::kLanes, bool full_queue = false) {
    Arena &a = ctx->arena;
    if (ctx->wide.base) {                // one traversal arena at a time: both are sized from the free memory
        NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->wide.base);
        ctx->wide = WideArena{};
    }
    const uint32_t max_levels = dist_bytes == 1 ? 254u : n + 1;
    ArenaPlan plan{};
    int prc = plan_arena(ctx, n, dist_bytes, want_slots, ksrc, plan, full_queue);
    if (prc != NEC_OK) return prc;
    const uint32_t slots = plan.slots;
    const size_t qcap = plan.qcap, ps = plan.per_slot;
    if (slots == 0) return fail(ctx, NEC_ENOMEM, "workspace budget %zu B cannot hold one batch slot (%zu B) for n=%u", plan.budget, ps, n);
    const bool reuse = a.base && a.n == n && a.dist_bytes == dist_bytes && a.ksrc == ksrc && a.n_slots >= slots && a.qcap == qcap;
    if (!reuse) {
        if (a.base) cudaFree(a.base);
        a = Arena{};
        const size_t bytes = ps * slots;
        cudaError_t e = cudaMalloc(&a.base, bytes);
        if (e != cudaSuccess) { a = Arena{}; return fail(ctx, NEC_ENOMEM, "cudaMalloc(%zu B workspace) failed: %s", bytes, cudaGetErrorString(e)); }
        a.bytes = bytes; a.n = n; a.n_slots = slots; a.dist_bytes = dist_bytes; a.qcap = qcap; a.max_levels = max_levels; a.ksrc = ksrc;
        // planes are grouped by kind so that one memset clears all slots' masks / partial loads
        uint8_t *p = (uint8_t *)a.base;
        auto carve = [&](size_t per_slot) { uint8_t *r = p; p += align_up(per_slot, 256) * slots; return r; };
        const size_t mask_b = ksrc / 8, qpitch = align_up(qcap, 64);
        a.dist = carve((size_t)n * ksrc * dist_bytes);              a.dist_stride = align_up((size_t)n * ksrc * dist_bytes, 256);
        a.q = (double *)carve((size_t)n * ksrc * 8);                a.q_stride = align_up((size_t)n * ksrc * 8, 256) / 8;
        a.mask = carve((size_t)n * 4 * mask_b);                     a.mask_stride = align_up((size_t)n * 4 * mask_b, 256);
        a.bslot = (double *)carve((size_t)n * 8);                   a.bslot_stride = align_up((size_t)n * 8, 256) / 8;
        a.queue_v = (uint32_t *)carve(qpitch * 4);                  a.queue_stride = qpitch;
        a.queue_m = carve(qpitch * mask_b);
        a.lvl = (uint32_t *)carve(((size_t)max_levels + 2) * 4);    a.lvl_stride = align_up(((size_t)max_levels + 2) * 4, 256) / 4;
    }
    return NEC_OK;
}

struct MeasureOut {                  // device outputs of the distance-measure mode, indexed like `sources`
    uint32_t *ecc;
    unsigned long long *sumdist;
    uint32_t *reached;
};

template <typename DistT, int kMode, int kThreads, int kSrc>
int launch_engine_t(nec_ctx *ctx, const nec::EngineParams &p, uint32_t slots) {
    constexpr size_t smem = nec::engine_smem_bytes<DistT, kThreads, kSrc>();
    // the attribute is per device (primary context), and contexts on several devices share this process: set it
    // on every launch (a host-side call of about a microsecond) instead of remembering it per instantiation
    if (smem > 48 * 1024)
        NEC_CUDA(ctx, cudaFuncSetAttribute(nec::vertex_load_engine<DistT, kMode, kThreads, kSrc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nec::vertex_load_engine<DistT, kMode, kThreads, kSrc><<<slots, kThreads, smem, ctx->stream>>>(p);
    NEC_CUDA(ctx, cudaGetLastError());
    return NEC_OK;
}

// Cluster teams: `team` CTAs of 512 threads per slot (grid = slots x team, cluster dimension = team).
template <typename DistT, int kSrc>
cudaLaunchConfig_t team_config(cudaLaunchAttribute *attr, uint32_t slots, uint32_t team, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(slots * team);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = nec::engine_smem_bytes<DistT, 512, kSrc>();
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}
// How many teams of `team` CTAs can be resident at once (0 on error).
template <typename DistT, int kSrc>
int team_capacity(uint32_t team) {
    auto kern = nec::vertex_load_engine<DistT, 0, 512, kSrc, true>;
    constexpr size_t smem = nec::engine_smem_bytes<DistT, 512, kSrc>();
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = team_config<DistT, kSrc>(attr, 1, team, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
template <typename DistT, int kSrc>
int launch_engine_team(nec_ctx *ctx, const nec::EngineParams &p, uint32_t slots, uint32_t team) {
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = team_config<DistT, kSrc>(attr, slots, team, ctx->stream);
    if (cfg.dynamicSmemBytes > 48 * 1024)      // per device, see launch_engine_t
        NEC_CUDA(ctx, cudaFuncSetAttribute(nec::vertex_load_engine<DistT, 0, 512, kSrc, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
    NEC_CUDA(ctx, cudaLaunchKernelEx(&cfg, nec::vertex_load_engine<DistT, 0, 512, kSrc, true>, p));
    return NEC_OK;
}

template <typename DistT, int kMode, int kSrc = 32>
int launch_engine(nec_ctx *ctx, const nec_graph *g, uint32_t n_sources, const uint32_t *d_worklist,
                  uint32_t n_work, int include_endpoints, uint32_t slots, uint32_t *dbg_npred, double *dbg_bk,
                  const MeasureOut *mo = nullptr) {
    double *dbg_sigma = ctx->dbg_sigma;
    const Arena &a = ctx->arena;
    nec::EngineParams p{};
    p.n = g->n; p.row_ptr = g->d_row_ptr; p.col_idx = g->d_col_idx;
    p.sources = ctx->d_sources; p.n_sources = n_sources;
    p.work_list = d_worklist; p.n_work = n_work; p.include_endpoints = include_endpoints;
    p.pull_mode = ctx->pull_mode;
    if (const char *env = getenv("NEC_PULL")) p.pull_mode = (uint32_t)std::min(2, std::max(0, atoi(env)));   // tests / A-B runs
    p.avg_deg = (uint32_t)std::max<uint64_t>(1, (g->nnz + g->n - 1) / std::max<uint32_t>(g->n, 1u));
    p.dist_base = a.dist; p.dist_stride = a.dist_stride;
    p.q_base = a.q; p.q_stride = a.q_stride;
    p.mask_base = a.mask; p.mask_stride = a.mask_stride;
    p.bslot_base = a.bslot; p.bslot_stride = a.bslot_stride;
    p.queue_v_base = a.queue_v; p.queue_m_base = a.queue_m; p.queue_stride = a.queue_stride; p.queue_cap = a.qcap;
    p.lvl_base = a.lvl; p.lvl_stride = a.lvl_stride;
    p.max_levels = a.max_levels;
    p.batch_status = ctx->d_status; p.counters = ctx->d_counters;
    p.dbg_npred = dbg_npred; p.dbg_bk = dbg_bk; p.dbg_sigma = dbg_npred ? dbg_sigma : nullptr;
    if (mo) { p.out_ecc = mo->ecc; p.out_sumdist = mo->sumdist; p.out_reached = mo->reached; }
    int rc;
    if constexpr (kMode == 0 && sizeof(DistT) == 1) {
        if (ctx->team_size > 1) rc = launch_engine_team<DistT, kSrc>(ctx, p, slots, ctx->team_size);
        else rc = ctx->wide_ctas ? launch_engine_t<DistT, kMode, 1024, kSrc>(ctx, p, slots) : launch_engine_t<DistT, kMode, 512, kSrc>(ctx, p, slots);
    } else
        rc = ctx->wide_ctas ? launch_engine_t<DistT, kMode, 1024, kSrc>(ctx, p, slots) : launch_engine_t<DistT, kMode, 512, kSrc>(ctx, p, slots);
    if (rc != NEC_OK) return rc;
    ctx->stats.kernel_launches++;
    return NEC_OK;
}

int engine_occupancy(nec_ctx *ctx) {
    int o8 = 0, o32 = 0;
    NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o8, nec::vertex_load_engine<uint8_t, 0, 512, 32>, 512, nec::engine_smem_bytes<uint8_t, 512, 32>()));
    NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o32, nec::vertex_load_engine<uint32_t, 0, 512, 32>, 512, nec::engine_smem_bytes<uint32_t, 512, 32>()));
    if (o8 < 1 || o32 < 1) return fail(ctx, NEC_ECUDA, "engine kernel cannot be resident (occupancy %d/%d)", o8, o32);
    ctx->occ_u8 = o8; ctx->occ_u32 = o32;
    for (auto &row : ctx->team_cap) for (int &c : row) c = -1;
    for (auto &a : ctx->wide_cap) for (auto &row : a) for (int &c : row) c = -1;
    return NEC_OK;
}

// How the batched engine runs a call: batch width, resident slots, CTA shape.  Shared by the run itself and
// by nec_wave_size.
struct BatchedPlan { uint32_t ksrc, want, slots, team; bool wide; };
int plan_batched(nec_ctx *ctx, uint32_t n, uint32_t n_sources, bool debug, bool measure, BatchedPlan &out) {
    const uint32_t sms = (uint32_t)ctx->prop.multiProcessorCount;
    // batch width: 64 sources per batch (the bit-parallel forward pass costs the same per adjacency entry at
    // either width, and a 64 B distance row is a whole DRAM burst) when the call has enough batches to fill the
    // slots a 64-wide layout gets AND those slots can keep the GPU busy: one per SM, or - a slot's state doubles
    // with the width, so huge graphs get fewer - at least the 2*SMs/8 that cluster teams of up to eight 512-thread
    // CTAs fill the GPU from.  Else 32: more, narrower batches.
    uint32_t ksrc = 32u;
    if (!debug && n_sources >= 64) {
        ArenaPlan plan64{};
        int prc = plan_arena(ctx, n, 1, sms, 64u, plan64);
        if (prc != NEC_OK) return prc;
        const uint32_t nb64 = (n_sources + 63) / 64, min_team_slots = (2 * sms + 7) / 8;
        if (plan64.slots >= sms ? nb64 >= sms : (plan64.slots >= min_team_slots && nb64 >= plan64.slots)) ksrc = 64u;
    }
    if (!debug && getenv("NEC_BATCH_WIDTH")) ksrc = atoi(getenv("NEC_BATCH_WIDTH")) == 64 ? 64u : 32u;
    const uint32_t n_batches = (n_sources + ksrc - 1) / ksrc;
    uint32_t want = std::min<uint32_t>(n_batches, (uint32_t)ctx->occ_u8 * sms);
    if (debug) want = 1;
    ArenaPlan ap{};
    int prc = plan_arena(ctx, n, 1, want, ksrc, ap);
    if (prc != NEC_OK) return prc;
    uint32_t slots = std::min<uint32_t>(want, ap.slots);
    // large graphs (beyond the mid engine) and memory-limited slot counts: 1024-thread CTAs, one per SM.
    // Same 32 warps per SM as 2 x 512, but half the resident slots (less HBM state in flight, smaller
    // level-boundary imbalance): 1.3x faster on the N = 2^20 graph in a same-box A/B.
    bool wide = !debug && (n > nec::kMidMaxN || (n_batches > slots && slots < sms + sms / 2));
    if (getenv("NEC_BATCHED_WIDE")) wide = atoi(getenv("NEC_BATCHED_WIDE")) != 0;
    if (wide) slots = std::min<uint32_t>(slots, sms);
    // fewer slots than SMs can keep busy (memory-limited: config 5): several 512-thread CTAs per slot as a
    // thread-block cluster.  Clusters are placed per GPC, so cudaOccupancyMaxActiveClusters decides how many
    // teams of each size fit (93 of 3, 71 of 4, 56 of 5, 45 of 6 on a B200); a batch finishes sooner the more
    // threads work on it (~ threads^0.75 measured), and the static batch -> slot mapping costs whole waves.
    uint32_t team_size = 1;
    if (wide && !measure && slots > 0 && n_batches >= slots) {
        const uint32_t tmax = std::min<uint32_t>(8u, 2u * sms / slots);
        const bool verbose = getenv("NEC_TEAM_VERBOSE") != nullptr;
        auto cost = [&](uint32_t s, double units) { return (double)((n_batches + s - 1) / s) / pow(units, 0.75); };
        double best = cost(slots, 2.0);                 // one 1024-thread CTA per slot = two 512-thread units
        uint32_t best_slots = slots;
        uint32_t t_lo = 3, t_hi = tmax;
        if (getenv("NEC_TEAM_SIZE")) { t_lo = t_hi = (uint32_t)atoi(getenv("NEC_TEAM_SIZE")); best = 1e300; }
        for (uint32_t team = t_hi; team >= t_lo && team >= 2 && team <= 8; --team) {
            int &cached = ctx->team_cap[ksrc == 64][team];
            if (cached < 0) cached = ksrc == 64 ? team_capacity<uint8_t, 64>(team) : team_capacity<uint8_t, 32>(team);
            const int cap = cached;
            if (verbose) fprintf(stderr, "[nec] cluster teams of %u: capacity %d, slots %u\n", team, cap, slots);
            if (cap <= 0) continue;
            const uint32_t s_t = std::min<uint32_t>(slots, (uint32_t)cap);
            const double c = cost(s_t, (double)team);
            if (c < best) { best = c; team_size = team; best_slots = s_t; }
        }
        slots = best_slots;
    }
    out = BatchedPlan{ksrc, want, slots, team_size, wide};
    return NEC_OK;
}

// Batched engine: partial load of `n_sources` sources into device vector d_out (n doubles).
int run_sources_batched(nec_ctx *ctx, const nec_graph *g, const uint32_t *h_sources, uint32_t n_sources,
                int include_endpoints, double *d_out, bool debug, uint32_t *dbg_npred, double *dbg_bk,
                const MeasureOut *mo = nullptr) {
    ctx->stats = nec_stats{};
    ctx->stats.n_sources = n_sources;
    const uint32_t n = g->n;
    if (n == 0) return NEC_OK;
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_sources == 0) {
        if (d_out) NEC_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * sizeof(double), ctx->stream));
        return NEC_OK;
    }
    for (uint32_t k = 0; k < n_sources; ++k)
        if (h_sources[k] >= n) return fail(ctx, NEC_EINVAL, "source %u (index %u) out of range for n=%u", h_sources[k], k, n);

    BatchedPlan plan{};
    int rc = plan_batched(ctx, n, n_sources, debug, mo != nullptr, plan);
    if (rc != NEC_OK) return rc;
    const uint32_t ksrc = plan.ksrc;
    const uint32_t sms = (uint32_t)ctx->prop.multiProcessorCount;
    const uint32_t n_batches32 = (n_sources + 31) / 32;
    const uint32_t n_batches = (n_sources + ksrc - 1) / ksrc;
    ctx->stats.n_batches = n_batches;
    ctx->stats.batch_width = ksrc;
    if ((rc = grow(ctx, ctx->d_sources, ctx->d_sources_cap, n_sources)) != NEC_OK) return rc;
    if ((rc = grow(ctx, ctx->d_status, ctx->d_status_cap, n_batches32)) != NEC_OK) return rc;
    NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_sources, h_sources, (size_t)n_sources * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, n_batches, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));

    // ---- pass 1: 8-bit distances -----------------------------------------------------
    if ((rc = ensure_arena(ctx, n, 1, plan.want, ksrc)) != NEC_OK) return rc;
    const uint32_t slots = std::min<uint32_t>(plan.slots, ctx->arena.n_slots);
    ctx->wide_ctas = plan.wide;
    ctx->team_size = slots == plan.slots ? plan.team : 1;
    const uint32_t team_used = ctx->team_size, cta_threads_used = ctx->team_size > 1 ? 512u : (ctx->wide_ctas ? 1024u : 512u);
    const Arena &a = ctx->arena;
    NEC_CUDA(ctx, cudaMemsetAsync(a.mask, 0, a.mask_stride * slots, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(a.bslot, 0, a.bslot_stride * sizeof(double) * slots, ctx->stream));
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (ksrc == 64)
        rc = mo ? launch_engine<uint8_t, 2, 64>(ctx, g, n_sources, nullptr, n_batches, include_endpoints, slots, nullptr, nullptr, mo)
                : launch_engine<uint8_t, 0, 64>(ctx, g, n_sources, nullptr, n_batches, include_endpoints, slots, nullptr, nullptr);
    else
        rc = mo    ? launch_engine<uint8_t, 2>(ctx, g, n_sources, nullptr, n_batches, include_endpoints, slots, nullptr, nullptr, mo)
           : debug ? launch_engine<uint8_t, 1>(ctx, g, n_sources, nullptr, n_batches, include_endpoints, slots, dbg_npred, dbg_bk)
                   : launch_engine<uint8_t, 0>(ctx, g, n_sources, nullptr, n_batches, include_endpoints, slots, nullptr, nullptr);
    if (rc != NEC_OK) return rc;
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev_mid, ctx->stream));
    if (!mo) {
        nec::reduce_slots_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(a.bslot, a.bslot_stride, slots, n, d_out);
        NEC_CUDA(ctx, cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->stats.n_slots = slots;
    ctx->stats.workspace_bytes = ctx->arena.bytes;

    // ---- read back the per-batch verdicts (small, but a host sync) ---------------------
    std::vector<uint8_t> status(n_batches);
    unsigned long long counters[8];
    NEC_CUDA(ctx, cudaMemcpyAsync(status.data(), ctx->d_status, n_batches, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaMemcpyAsync(counters, ctx->d_counters, sizeof counters, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.kernel_ms = ms;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_mid));
    ctx->stats.engine_ms = ms;

    // Batches that did not complete are re-run, never approximated: a queue overflow (the planner halves the
    // queue bound under memory pressure to keep one slot per SM) with the full n * width queue on fewer slots, a
    // BFS deeper than 253 levels with 32-bit distance rows.  `rerun` keeps the sums so far, runs the listed batch
    // ids (of the given width) in a freshly planned arena and reads their verdicts back.
    auto rerun = [&](const std::vector<uint32_t> &list, uint32_t dist_bytes, uint32_t width, bool full_queue,
                     std::vector<uint8_t> &verdict) -> int {
        int r;
        if ((r = grow(ctx, ctx->d_keep, ctx->d_keep_cap, n)) != NEC_OK) return r;
        if (!mo) NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_keep, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if ((r = grow(ctx, ctx->d_worklist, ctx->d_worklist_cap, list.size())) != NEC_OK) return r;
        NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_worklist, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        const uint32_t n_ids = (n_sources + width - 1) / width;
        NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, n_ids, ctx->stream));
        uint32_t want2 = std::min<uint32_t>((uint32_t)list.size(), (uint32_t)((dist_bytes == 1 ? ctx->occ_u8 : ctx->occ_u32) * ctx->prop.multiProcessorCount));
        if (debug) want2 = 1;
        ctx->wide_ctas = false;
        ctx->team_size = 1;
        NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the arena may be reallocated below
        if ((r = ensure_arena(ctx, n, dist_bytes, want2, width, full_queue)) != NEC_OK) return r;
        const Arena &w = ctx->arena;
        const uint32_t slots2 = std::min<uint32_t>(want2, w.n_slots);
        NEC_CUDA(ctx, cudaMemsetAsync(w.mask, 0, w.mask_stride * slots2, ctx->stream));
        NEC_CUDA(ctx, cudaMemsetAsync(w.bslot, 0, w.bslot_stride * sizeof(double) * slots2, ctx->stream));
        NEC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        const uint32_t cnt = (uint32_t)list.size();
        if (dist_bytes == 4)
            r = mo    ? launch_engine<uint32_t, 2>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr, mo)
              : debug ? launch_engine<uint32_t, 1>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, dbg_npred, dbg_bk)
                      : launch_engine<uint32_t, 0>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr);
        else if (width == 64)
            r = mo ? launch_engine<uint8_t, 2, 64>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr, mo)
                   : launch_engine<uint8_t, 0, 64>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr);
        else
            r = mo    ? launch_engine<uint8_t, 2>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr, mo)
              : debug ? launch_engine<uint8_t, 1>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, dbg_npred, dbg_bk)
                      : launch_engine<uint8_t, 0>(ctx, g, n_sources, ctx->d_worklist, cnt, include_endpoints, slots2, nullptr, nullptr);
        if (r != NEC_OK) return r;
        NEC_CUDA(ctx, cudaEventRecord(ctx->ev_mid, ctx->stream));
        if (!mo) {
            // slot 0's partial absorbs the sums so far so that one reduction produces the result
            nec::add_into_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(w.bslot, ctx->d_keep, n);
            nec::reduce_slots_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(w.bslot, w.bslot_stride, slots2, n, d_out);
            NEC_CUDA(ctx, cudaGetLastError());
            ctx->stats.kernel_launches += 2;
        }
        NEC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        verdict.assign(n_ids, 0);
        NEC_CUDA(ctx, cudaMemcpyAsync(verdict.data(), ctx->d_status, n_ids, cudaMemcpyDeviceToHost, ctx->stream));
        NEC_CUDA(ctx, cudaMemcpyAsync(counters, ctx->d_counters, sizeof counters, cudaMemcpyDeviceToHost, ctx->stream));
        NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms2 = 0.f;
        NEC_CUDA(ctx, cudaEventElapsedTime(&ms2, ctx->ev0, ctx->ev1));
        ctx->stats.kernel_ms += ms2;
        NEC_CUDA(ctx, cudaEventElapsedTime(&ms2, ctx->ev0, ctx->ev_mid));
        ctx->stats.engine_ms += ms2;
        ctx->stats.workspace_bytes = std::max<uint64_t>(ctx->stats.workspace_bytes, ctx->arena.bytes);
        return NEC_OK;
    };

    std::vector<uint32_t> deep;                     // 32-source batch ids for the 32-bit distance pass
    std::vector<uint32_t> overflowed;               // batch ids (of width ksrc) whose level queue overflowed
    auto sort_verdicts = [&](const std::vector<uint8_t> &verdict, const uint32_t *ids, uint32_t count, bool allow_overflow) -> int {
        for (uint32_t i = 0; i < count; ++i) {
            const uint32_t b = ids ? ids[i] : i;
            if (verdict[b] == nec::kTooDeep) {
                if (ksrc == 32) deep.push_back(b);
                else { deep.push_back(2 * b); if (2 * b + 1 < n_batches32) deep.push_back(2 * b + 1); }
            } else if (verdict[b] == nec::kQueueOverflow) {
                if (!allow_overflow)
                    return fail(ctx, NEC_EINTERNAL, "frontier queue overflow in batch %u (capacity %zu entries); raise the workspace limit", b, ctx->arena.qcap);
                overflowed.push_back(b);
            } else if (verdict[b] != nec::kDone)
                return fail(ctx, NEC_EINTERNAL, "batch %u was not processed (status %u)", b, (unsigned)verdict[b]);
        }
        return NEC_OK;
    };
    const bool queue_was_cut = ctx->arena.qcap < (size_t)n * ksrc + ksrc;
    if ((rc = sort_verdicts(status, nullptr, n_batches, queue_was_cut)) != NEC_OK) return rc;

    if (!overflowed.empty()) {
        ArenaPlan full{};
        if ((rc = plan_arena(ctx, n, 1, 1, ksrc, full, true)) != NEC_OK) return rc;
        if (full.slots == 0)
            return fail(ctx, NEC_EINTERNAL, "frontier queue overflow in batch %u (capacity %zu entries) and the workspace limit cannot hold one slot with "
                        "the full queue (%zu B); raise the workspace limit", overflowed[0], ctx->arena.qcap, full.per_slot);
        ctx->stats.requeued_batches = (uint32_t)overflowed.size();
        const std::vector<uint32_t> list = overflowed;
        overflowed.clear();
        std::vector<uint8_t> verdict;
        if ((rc = rerun(list, 1, ksrc, true, verdict)) != NEC_OK) return rc;
        if ((rc = sort_verdicts(verdict, list.data(), (uint32_t)list.size(), false)) != NEC_OK) return rc;
    }

    // ---- batches deeper than 253 levels: 32-bit distances, 32 sources per batch ----------------
    if (!deep.empty()) {
        ctx->stats.wide_batches = (uint32_t)deep.size();
        std::vector<uint8_t> verdict;
        if ((rc = rerun(deep, 4, 32, false, verdict)) != NEC_OK) return rc;
        for (uint32_t b : deep) {
            if (verdict[b] == nec::kQueueOverflow)
                return fail(ctx, NEC_EINTERNAL, "frontier queue overflow in batch %u (32-bit distance pass, capacity %zu entries); raise the workspace limit", b, ctx->arena.qcap);
            if (verdict[b] != nec::kDone)
                return fail(ctx, NEC_EINTERNAL, "batch %u failed in the 32-bit distance pass (status %u)", b, (unsigned)verdict[b]);
        }
    }
    ctx->stats.engine = 1;
    ctx->stats.team_size = team_used;
    ctx->stats.cta_threads = cta_threads_used;
    ctx->stats.pull_levels = (uint32_t)(counters[7] >> 32);
    ctx->stats.push_levels = (uint32_t)counters[7];
    ctx->stats.reached_pairs = counters[0];
    ctx->stats.edge_pairs = counters[1];
    ctx->stats.vertex_visits = counters[2];
    ctx->stats.max_depth = (uint32_t)counters[3];
    {
        const double tot = (double)(counters[4] + counters[5] + counters[6]);
        ctx->stats.frac_reset = tot > 0 ? (float)(counters[4] / tot) : 0.f;
        ctx->stats.frac_forward = tot > 0 ? (float)(counters[5] / tot) : 0.f;
        ctx->stats.frac_backward = tot > 0 ? (float)(counters[6] / tot) : 0.f;
    }
    return NEC_OK;
}

// ---- compact (wide) engine, nec_wide.cuh ------------------------------------------------------------------------
#ifndef NEC_WIDE_COST256
#define NEC_WIDE_COST256 2.6         // planner: time of a 256-wide batch relative to a 64-wide one on the same threads (128-wide: 1.5);
                                     // measured on config 4: 484 ms on 3 CTAs against 414 ms on 2 CTAs for a 128-wide batch
#endif
#ifndef NEC_WIDE_SMALL_CTA
#define NEC_WIDE_SMALL_CTA 512       // the alternative CTA shape (NEC_WIDE_THREADS selects it): 512 threads, two CTAs per SM
#endif
size_t wide_per_slot_bytes(uint32_t n, uint32_t kw, size_t qcap, uint32_t max_levels) {
    const size_t qpitch = align_up(qcap, 64);
    size_t b = 0;
    b += align_up((size_t)n * (kw / 2), 256);                       // forward records
    b += align_up((size_t)n * kw, 256);                             // level records: 4 slots of kw / 4 bytes
    b += align_up(((size_t)n * kw + 64) * sizeof(double), 256);     // compact q runs
    b += align_up((size_t)n * sizeof(double), 256) * (kw == 256 ? 2 : 1);   // private partial load (256 per batch: one per half of an entry)
    b += qpitch * 4 * 2 + qpitch * (kw / 8);                        // queue: vertex, position, mask
    b += align_up(((size_t)max_levels + 2) * sizeof(uint32_t), 256);
    if (kw == 256) b += 2 * qpitch * 4 + align_up(((size_t)max_levels + 2) * sizeof(uint32_t), 256);   // backward work list, per-level counts
    return b;
}

template <int kW, int kThreads>
cudaLaunchConfig_t wide_config(cudaLaunchAttribute *attr, uint32_t slots, uint32_t team, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(slots * team);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = nec::wide_smem_bytes<kW, kThreads>();
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}
template <int kW, int kThreads>
int wide_capacity_t(uint32_t team) {
    auto kern = nec::wide_engine<kW, kThreads>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nec::wide_smem_bytes<kW, kThreads>()) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = wide_config<kW, kThreads>(attr, 1, team, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int wide_capacity(nec_ctx *ctx, uint32_t kw, uint32_t threads, uint32_t team) {
    int &cached = ctx->wide_cap[kw == 64 ? 0 : kw == 128 ? 1 : 2][threads == 1024][team];
    if (cached < 0)
        cached = kw == 256 ? (threads == 1024 ? wide_capacity_t<256, 1024>(team) : wide_capacity_t<256, NEC_WIDE_SMALL_CTA>(team))
               : kw == 128 ? (threads == 1024 ? wide_capacity_t<128, 1024>(team) : wide_capacity_t<128, NEC_WIDE_SMALL_CTA>(team))
                           : (threads == 1024 ? wide_capacity_t<64, 1024>(team) : wide_capacity_t<64, NEC_WIDE_SMALL_CTA>(team));
    return cached;
}
// Software teams: any `team` consecutive CTAs of a cooperative launch; what limits them is only how many CTAs are co-resident.
template <int kW, int kThreads>
int wide_soft_ctas_t(int sms) {
    auto kern = nec::wide_engine<kW, kThreads, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nec::wide_smem_bytes<kW, kThreads>()) != cudaSuccess) { cudaGetLastError(); return 0; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, nec::wide_smem_bytes<kW, kThreads>()) != cudaSuccess) { cudaGetLastError(); return 0; }
    return occ * sms;
}
int wide_soft_ctas(nec_ctx *ctx, uint32_t kw, uint32_t threads) {
    int &cached = ctx->wide_cap[kw == 64 ? 0 : kw == 128 ? 1 : 2][threads == 1024][0];      // team index 0: resident CTAs of the software-team kernel
    const int sms = ctx->prop.multiProcessorCount;
    if (cached < 0)
        cached = kw == 256 ? (threads == 1024 ? wide_soft_ctas_t<256, 1024>(sms) : wide_soft_ctas_t<256, NEC_WIDE_SMALL_CTA>(sms))
               : kw == 128 ? (threads == 1024 ? wide_soft_ctas_t<128, 1024>(sms) : wide_soft_ctas_t<128, NEC_WIDE_SMALL_CTA>(sms))
                           : (threads == 1024 ? wide_soft_ctas_t<64, 1024>(sms) : wide_soft_ctas_t<64, NEC_WIDE_SMALL_CTA>(sms));
    return cached;
}
template <int kW, int kThreads>
int launch_wide_soft_t(nec_ctx *ctx, const nec::WideParams &p, uint32_t slots, uint32_t team) {
    auto kern = nec::wide_engine<kW, kThreads, true>;
    NEC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nec::wide_smem_bytes<kW, kThreads>()));   // per device
    if (const char *env = getenv("NEC_WIDE_CARVEOUT")) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(env));   // A/B knob (percent)
    void *args[] = {(void *)&p};
    NEC_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kern, dim3(slots * team), dim3(kThreads), args, nec::wide_smem_bytes<kW, kThreads>(), ctx->stream));
    return NEC_OK;
}

template <int kW, int kThreads>
int launch_wide_t(nec_ctx *ctx, const nec::WideParams &p, uint32_t slots, uint32_t team) {
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = wide_config<kW, kThreads>(attr, slots, team, ctx->stream);
    NEC_CUDA(ctx, cudaFuncSetAttribute(nec::wide_engine<kW, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));   // per device
    NEC_CUDA(ctx, cudaLaunchKernelEx(&cfg, nec::wide_engine<kW, kThreads>, p));
    return NEC_OK;
}

// How the wide engine runs a call: batch width, CTA shape, resident slots and CTAs per slot.
struct WidePlan { uint32_t kw, threads, slots, team, max_levels; size_t qcap, per_slot; bool soft; };
int plan_wide(nec_ctx *ctx, uint32_t n, uint32_t n_sources, WidePlan &out) {
    const uint32_t sms = (uint32_t)ctx->prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    NEC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = ctx->ws_limit ? (size_t)ctx->ws_limit : (size_t)((double)(free_b + ctx->arena.bytes + ctx->wide.bytes) * 0.8);
    uint32_t threads = 1024;
    if (const char *env = getenv("NEC_WIDE_THREADS")) threads = atoi(env) == 1024 ? 1024u : (uint32_t)NEC_WIDE_SMALL_CTA;
    const uint32_t ctas_total = threads <= 512 ? 2 * sms : sms;
    const uint32_t max_levels = n;
    // queue entries per vertex the slot has room for: a vertex sits on a handful of distinct levels within one batch
    // on the graphs this engine is meant for; batches that need more are handed to the source-minor engine
    // A dozen suffices on random graphs (5 at 256 per batch); small-world graphs need ~14 (half the batches of the N=2^16 config overflow a
    // queue of 12: 28 against 53 GTEPS).  Room for 24 is planned wherever it costs no speed (with software teams fewer, larger teams run
    // as fast as many small ones), 12 where memory is so short that it would.
    size_t per_vertex_opts[2] = {24, 12};
    int n_opts = 2;
    if (const char *env = getenv("NEC_WIDE_QUEUE")) { per_vertex_opts[0] = (size_t)std::max(1, atoi(env)); n_opts = 1; }
    WidePlan best{};
    double best_cost = 1e300;
    for (int opt = 0; opt < n_opts; ++opt)
    for (uint32_t kw : {256u, 128u, 64u}) {
        const size_t per_vertex = per_vertex_opts[opt];
        if (ctx->probe_mode) { if (kw != 64u) continue; }               // engine_hint's probe: four 64-source batches, whatever the knobs say
        else if (const char *env = getenv("NEC_WIDE_WIDTH")) { if ((uint32_t)atoi(env) != kw) continue; }
        if ((uint64_t)n * kw + kw >= 0xFFFFFFFFull) continue;           // 32-bit positions and queue indices
        const size_t qcap = std::min((size_t)n * kw + kw, (size_t)n * per_vertex + kw);
        const size_t ps = wide_per_slot_bytes(n, kw, qcap, max_levels);
        const uint32_t n_batches = (n_sources + kw - 1) / kw;
        const uint32_t fit = (uint32_t)std::min<size_t>(n_batches, budget / ps);
        if (fit == 0) continue;
        // every team size the slots allow: `team` CTAs per slot fill the GPU from ctas_total / team slots (clusters are
        // placed per GPC, so the driver is asked how many of each size can be resident)
        // teams: thread-block clusters (placed per GPC: the driver is asked how many of each size can be resident), or
        // software teams over a cooperative launch (any size, every SM usable) - the default
        bool soft = true;
        if (const char *env = getenv("NEC_WIDE_SOFT")) soft = atoi(env) != 0;
        auto capacity = [&](uint32_t team) -> int { return soft ? wide_soft_ctas(ctx, kw, threads) / (int)team : wide_capacity(ctx, kw, threads, team); };
        uint32_t t_lo = 1, t_hi = soft ? 16 : 8;                         // (a portable cluster has at most 8 CTAs)
        if (const char *env = getenv("NEC_WIDE_TEAM")) {                 // tests / A-B runs: this team size, or the largest below it that fits
            t_hi = (uint32_t)std::min((int)t_hi, std::max(1, atoi(env)));
            while (t_hi > 1 && capacity(t_hi) <= 0) --t_hi;
            t_lo = t_hi;
        }
        for (uint32_t team = t_lo; team <= t_hi; ++team) {
            const int cap = capacity(team);
            if (cap <= 0) continue;
            const uint32_t slots = std::min<uint32_t>(std::min<uint32_t>(fit, (uint32_t)cap), std::max<uint32_t>(1u, ctas_total / team));
            // relative cost: waves x time of one batch.  Measured (scripts/exp_wide_teams.py): a batch's time is
            // inversely proportional to the threads working on it (12 batches alone on the GPU: 2321 / 1178 / 809 /
            // 615 / 429 / 337 ms with 1 / 2 / 3 / 4 / 6 / 8 CTAs each), with the GPU full the throughput does not depend
            // on the team size, and a 128-wide batch costs 1.5 x a 64-wide one, a 256-wide one 2.6 x (the forward pass and
            // the row traffic are per batch).  Equal costs keep the smaller team: more slots, finer last wave.
            const double waves = (double)((n_batches + slots - 1) / slots);
            const double cost = waves * (kw == 256 ? NEC_WIDE_COST256 : kw == 128 ? 1.5 : 1.0) / ((double)team * threads / 512.0);
            if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = WidePlan{kw, threads, slots, team, max_levels, qcap, ps, soft}; }
        }
    }
    if (best.slots == 0) return fail(ctx, NEC_ENOMEM, "workspace budget %zu B cannot hold one wide-engine slot for n=%u", budget, n);
    out = best;
    return NEC_OK;
}

int ensure_wide_arena(nec_ctx *ctx, uint32_t n, const WidePlan &pl) {
    WideArena &a = ctx->wide;
    if (ctx->arena.base) {                  // one traversal arena at a time
        NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->arena.base);
        ctx->arena = Arena{};
    }
    if (a.base && a.n == n && a.kw == pl.kw && a.n_slots >= pl.slots && a.qcap == pl.qcap && a.max_levels == pl.max_levels) return NEC_OK;
    if (a.base) { NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(a.base); }
    a = WideArena{};
    const uint32_t slots = pl.slots, kw = pl.kw;
    const size_t bytes = pl.per_slot * slots;
    cudaError_t e = cudaMalloc(&a.base, bytes);
    if (e != cudaSuccess) { a = WideArena{}; return fail(ctx, NEC_ENOMEM, "cudaMalloc(%zu B wide-engine workspace) failed: %s", bytes, cudaGetErrorString(e)); }
    a.bytes = bytes; a.n = n; a.n_slots = slots; a.kw = kw; a.qcap = pl.qcap; a.max_levels = pl.max_levels;
    uint8_t *p = (uint8_t *)a.base;
    auto carve = [&](size_t per_slot) { uint8_t *r = p; p += align_up(per_slot, 256) * slots; return r; };
    const size_t qpitch = align_up(pl.qcap, 64);
    a.rec = carve((size_t)n * (kw / 2));                           a.rec_stride = align_up((size_t)n * (kw / 2), 256);
    a.lr = carve((size_t)n * kw);                                  a.lr_stride = align_up((size_t)n * kw, 256);
    a.qc = (double *)carve(((size_t)n * kw + 64) * 8);             a.qc_stride = align_up(((size_t)n * kw + 64) * 8, 256) / 8;
    const size_t bhalf = align_up((size_t)n * 8, 256);             // 256 per batch: two partial loads per slot (first / second half of an entry)
    a.bslot = (double *)carve(kw == 256 ? 2 * bhalf : (size_t)n * 8); a.bslot_stride = (kw == 256 ? 2 * bhalf : bhalf) / 8;
    a.queue_v = (uint32_t *)carve(qpitch * 4);                     a.queue_stride = qpitch;
    a.queue_p = (uint32_t *)carve(qpitch * 4);
    a.queue_m = carve(qpitch * (kw / 8));
    a.lvl = (uint32_t *)carve(((size_t)pl.max_levels + 2) * 4);    a.lvl_stride = align_up(((size_t)pl.max_levels + 2) * 4, 256) / 4;
    if (kw == 256) {
        a.bq = (uint32_t *)carve(2 * qpitch * 4);                  a.bq_stride = 2 * qpitch;
        a.bcnt = (uint32_t *)carve(((size_t)pl.max_levels + 2) * 4);
    }
    return NEC_OK;
}

// Wide engine: partial load of `n_sources` sources into device vector d_out (n doubles).  Batches it cannot hold
// (level queue beyond the slot's room) are re-run by the source-minor engine, never approximated.
int run_sources_wide(nec_ctx *ctx, const nec_graph *g, const uint32_t *h_sources, uint32_t n_sources, int include_endpoints, double *d_out) {
    ctx->stats = nec_stats{};
    ctx->stats.n_sources = n_sources;
    const uint32_t n = g->n;
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    WidePlan pl{};
    int rc = plan_wide(ctx, n, n_sources, pl);
    if (rc != NEC_OK) return rc;
    const uint32_t kw = pl.kw, n_batches = (n_sources + kw - 1) / kw;
    if ((rc = grow(ctx, ctx->d_sources, ctx->d_sources_cap, n_sources)) != NEC_OK) return rc;
    if ((rc = grow(ctx, ctx->d_status, ctx->d_status_cap, n_batches)) != NEC_OK) return rc;
    if ((rc = ensure_wide_arena(ctx, n, pl)) != NEC_OK) return rc;
    const WideArena &a = ctx->wide;
    const uint32_t slots = pl.slots;
    NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_sources, h_sources, (size_t)n_sources * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, n_batches, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(a.bslot, 0, a.bslot_stride * sizeof(double) * slots, ctx->stream));
    nec::WideParams p{};
    p.n = n; p.row_ptr = g->d_row_ptr; p.col_idx = g->d_col_idx;
    p.sources = ctx->d_sources; p.n_sources = n_sources; p.work_list = nullptr; p.n_work = n_batches;
    p.include_endpoints = include_endpoints;
    p.pull_mode = ctx->pull_mode;
    if (const char *env = getenv("NEC_PULL")) p.pull_mode = (uint32_t)std::min(2, std::max(0, atoi(env)));
    p.avg_deg = (uint32_t)std::max<uint64_t>(1, (g->nnz + g->n - 1) / std::max<uint32_t>(g->n, 1u));
    p.rec_base = a.rec; p.rec_stride = a.rec_stride;
    p.lr_base = a.lr; p.lr_stride = a.lr_stride;
    p.qc_base = a.qc; p.qc_stride = a.qc_stride;
    p.bslot_base = a.bslot; p.bslot_stride = a.bslot_stride;
    p.queue_v_base = a.queue_v; p.queue_m_base = a.queue_m; p.queue_p_base = a.queue_p; p.queue_stride = a.queue_stride; p.queue_cap = a.qcap;
    p.lvl_base = a.lvl; p.lvl_stride = a.lvl_stride; p.max_levels = a.max_levels;
    p.bq_base = a.bq; p.bq_stride = a.bq_stride; p.bcnt_base = a.bcnt; p.bslot_hi_off = a.bslot_stride / 2;
    p.batch_status = ctx->d_status; p.counters = ctx->d_counters;
    p.team_size = pl.team; p.team_base = nullptr;
    p.forward_only = ctx->probe_mode ? 1 : 0;
    if (pl.soft) {                                   // the teams' shared scalars and barrier counters (8 u32 per slot)
        if ((rc = grow(ctx, ctx->d_worklist, ctx->d_worklist_cap, (size_t)slots * 8)) != NEC_OK) return rc;
        NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_worklist, 0, (size_t)slots * 8 * sizeof(uint32_t), ctx->stream));
        p.team_base = ctx->d_worklist;
    }
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (pl.soft) {
        if (kw == 256) rc = pl.threads == 1024 ? launch_wide_soft_t<256, 1024>(ctx, p, slots, pl.team) : launch_wide_soft_t<256, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
        else if (kw == 128) rc = pl.threads == 1024 ? launch_wide_soft_t<128, 1024>(ctx, p, slots, pl.team) : launch_wide_soft_t<128, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
        else rc = pl.threads == 1024 ? launch_wide_soft_t<64, 1024>(ctx, p, slots, pl.team) : launch_wide_soft_t<64, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
    } else {
        if (kw == 256) rc = pl.threads == 1024 ? launch_wide_t<256, 1024>(ctx, p, slots, pl.team) : launch_wide_t<256, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
        else if (kw == 128) rc = pl.threads == 1024 ? launch_wide_t<128, 1024>(ctx, p, slots, pl.team) : launch_wide_t<128, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
        else rc = pl.threads == 1024 ? launch_wide_t<64, 1024>(ctx, p, slots, pl.team) : launch_wide_t<64, NEC_WIDE_SMALL_CTA>(ctx, p, slots, pl.team);
    }
    if (rc != NEC_OK) return rc;
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev_mid, ctx->stream));
    if (kw == 256)    // two partial loads per slot, summed in (slot, half) order
        nec::reduce_slots_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(a.bslot, a.bslot_stride / 2, 2 * slots, n, d_out);
    else
        nec::reduce_slots_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(a.bslot, a.bslot_stride, slots, n, d_out);
    NEC_CUDA(ctx, cudaGetLastError());
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<uint8_t> status(n_batches);
    unsigned long long counters[8];
    NEC_CUDA(ctx, cudaMemcpyAsync(status.data(), ctx->d_status, n_batches, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaMemcpyAsync(counters, ctx->d_counters, sizeof counters, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    nec_stats st{};
    st.n_sources = n_sources; st.n_batches = n_batches; st.batch_width = kw; st.n_slots = slots; st.team_size = pl.team; st.cta_threads = pl.threads;
    st.workspace_bytes = a.bytes; st.kernel_launches = 2; st.engine = 4;
    float ms = 0.f;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    st.kernel_ms = ms;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_mid));
    st.engine_ms = ms;
    st.pull_levels = (uint32_t)(counters[7] >> 32); st.push_levels = (uint32_t)counters[7];
    st.reached_pairs = counters[0]; st.edge_pairs = counters[1]; st.vertex_visits = counters[2]; st.max_depth = (uint32_t)counters[3];
    {
        const double tot = (double)(counters[4] + counters[5] + counters[6]);
        st.frac_reset = tot > 0 ? (float)(counters[4] / tot) : 0.f;
        st.frac_forward = tot > 0 ? (float)(counters[5] / tot) : 0.f;
        st.frac_backward = tot > 0 ? (float)(counters[6] / tot) : 0.f;
    }
    std::vector<uint32_t> redo;                     // sources of the batches the slot could not hold
    for (uint32_t b = 0; b < n_batches; ++b) {
        if (status[b] == nec::kDone) continue;
        if (status[b] != nec::kQueueOverflow && status[b] != nec::kTooDeep)
            return fail(ctx, NEC_EINTERNAL, "wide engine: batch %u was not processed (status %u)", b, (unsigned)status[b]);
        for (uint32_t k = b * kw; k < std::min<uint64_t>((uint64_t)(b + 1) * kw, n_sources); ++k) redo.push_back(h_sources[k]);
        st.requeued_batches++;
    }
    if (!redo.empty() && !ctx->probe_mode) {            // (a probe only wants to know that the batch did not fit)
        if ((rc = grow(ctx, ctx->d_tmp, ctx->d_tmp_cap, n)) != NEC_OK) return rc;
        rc = run_sources_batched(ctx, g, redo.data(), (uint32_t)redo.size(), include_endpoints, ctx->d_tmp, false, nullptr, nullptr, nullptr);
        if (rc != NEC_OK) return rc;
        nec::add_into_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(d_out, ctx->d_tmp, n);
        NEC_CUDA(ctx, cudaGetLastError());
        const nec_stats second = ctx->stats;
        st.reached_pairs += second.reached_pairs; st.edge_pairs += second.edge_pairs; st.vertex_visits += second.vertex_visits;
        st.max_depth = std::max(st.max_depth, second.max_depth);
        st.kernel_launches += second.kernel_launches + 1;
        st.kernel_ms += second.kernel_ms; st.engine_ms += second.engine_ms;
        st.wide_batches += second.wide_batches;
        st.engine = 4u | (1u << 4);
        st.workspace_bytes = std::max(st.workspace_bytes, second.workspace_bytes);
    }
    ctx->stats = st;
    return NEC_OK;
}

// One-source-per-CTA engine (nec_mid.cuh).  Sources it hands back (too deep / predecessor
// count overflow) are appended to `fallback`.
int run_sources_mid(nec_ctx *ctx, const nec_graph *g, const uint32_t *h_sources, uint32_t n_sources,
                    int include_endpoints, double *d_out, std::vector<uint32_t> &fallback,
                    const MeasureOut *mo = nullptr, std::vector<uint32_t> *fallback_index = nullptr) {
    const uint32_t n = g->n;
    ctx->stats = nec_stats{};
    ctx->stats.n_sources = n_sources;
    ctx->stats.n_batches = n_sources;
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    // CTA shape: 2 x 512 threads per SM, or - small graphs, whose levels hold a few hundred vertices - 8 x 128:
    // a source's time is a chain of per-level latencies there, so sources in flight matter, not threads per source
    // (same-box A/B on small-world graphs: 1.94x at n = 2048, 1.74x at 4096, 1.50x at 8192, 1.14x at 16384)
    uint32_t small_n = 16384;
    if (const char *env = getenv("NEC_MID_SMALL_N")) small_n = (uint32_t)atol(env);
    const bool small_ctas = n <= small_n;
    const int threads = small_ctas ? nec::kMidThreadsSmall : nec::kMidThreads;
    const size_t smem = nec::mid_smem_bytes(n, g->slab_w, threads);
    int occ = 0;
    if (ctx->mid_occ.occ > 0 && ctx->mid_occ.n == n && ctx->mid_occ.slab_w == g->slab_w && ctx->mid_occ.threads == threads) occ = ctx->mid_occ.occ;
    else {
        // the attribute is per device; this context is bound to one device, so once per context (and shape) suffices
        const int big = (int)(227 * 1024 - 4096);
        NEC_CUDA(ctx, cudaFuncSetAttribute(nec::mid_engine<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        NEC_CUDA(ctx, cudaFuncSetAttribute(nec::mid_engine<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        if (small_ctas) {
            if (g->slab_w == 8) NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nec::mid_engine<8, nec::kMidThreadsSmall>, threads, smem));
            else NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nec::mid_engine<16, nec::kMidThreadsSmall>, threads, smem));
        } else {
            if (g->slab_w == 8) NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nec::mid_engine<8>, threads, smem));
            else NEC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nec::mid_engine<16>, threads, smem));
        }
        ctx->mid_occ.n = n; ctx->mid_occ.slab_w = g->slab_w; ctx->mid_occ.threads = threads; ctx->mid_occ.occ = occ;
    }
    if (occ < 1) return fail(ctx, NEC_ECUDA, "mid engine cannot be resident for n=%u (%zu B shared memory)", n, smem);
    if (const char *env = getenv("NEC_MID_CTAS_PER_SM")) {       // profiling knob
        const int v = atoi(env);
        if (v >= 1 && v < occ) occ = v;
    }
    uint32_t slots = std::min<uint32_t>(n_sources, (uint32_t)(occ * ctx->prop.multiProcessorCount));
    if (const char *env = getenv("NEC_MID_SLOTS")) {             // profiling knob: total resident CTAs
        const int v = atoi(env);
        if (v >= 1 && (uint32_t)v < slots) slots = (uint32_t)v;
    }
    const size_t q_stride = align_up((size_t)n * 8, 256) / 8;
    size_t o_stride = align_up((size_t)n * 2 * 4, 256) / 4;   // queue: room for duplicates
    if (const char *env = getenv("NEC_MID_QUEUE_CAP")) {          // test hook: force the overflow hand-back path
        const long v = atol(env);
        if (v >= 64 && (size_t)v < o_stride) o_stride = align_up((size_t)v * 4, 256) / 4;
    }
    const size_t need = (q_stride * 8 * 2 + o_stride * 4 + o_stride * 16) * slots;
    if (need > ctx->mid_bytes) {
        if (ctx->mid_base) cudaFree(ctx->mid_base);
        ctx->mid_base = nullptr; ctx->mid_bytes = 0;
        cudaError_t e = cudaMalloc(&ctx->mid_base, need);
        if (e != cudaSuccess) return fail(ctx, NEC_ENOMEM, "cudaMalloc(%zu B mid-engine workspace) failed: %s", need, cudaGetErrorString(e));
        ctx->mid_bytes = need;
    }
    int rc;
    if ((rc = grow(ctx, ctx->d_sources, ctx->d_sources_cap, n_sources)) != NEC_OK) return rc;
    // device: [8 counters | status per source] in one buffer (one memset, one copy back); host: pinned staging for the
    // source list and for that buffer, so that every transfer of the call is asynchronous and the host waits once
    const size_t ret_bytes = 64 + (size_t)n_sources;
    if ((rc = grow(ctx, ctx->d_src_status, ctx->d_src_status_cap, ret_bytes)) != NEC_OK) return rc;
    const size_t src_bytes = align_up((size_t)n_sources * 4, 64);
    if ((rc = grow_pinned(ctx, src_bytes + ret_bytes)) != NEC_OK) return rc;
    uint8_t *h_ret = ctx->h_pin + src_bytes;
    unsigned long long *d_counters = reinterpret_cast<unsigned long long *>(ctx->d_src_status);
    uint8_t *d_status = ctx->d_src_status + 64;
    double *d_q = (double *)ctx->mid_base;
    double *d_bslot = d_q + q_stride * slots;
    uint32_t *d_order = (uint32_t *)(d_bslot + q_stride * slots);
    uint4 *d_pos4 = (uint4 *)(d_order + o_stride * slots);   // o_stride entries per slot; 256 B aligned since o_stride*4 is
    const bool timing = getenv("NEC_MID_TIMING") != nullptr;
    const auto t_host0 = std::chrono::steady_clock::now();
    // a run of consecutive vertex ids (every call of nec_vertex_load, whole-graph shards) needs no source list on the device
    bool consecutive = n_sources <= 65536;
    for (uint32_t k = 1; consecutive && k < n_sources; ++k) consecutive = h_sources[k] == h_sources[0] + k;
    if (!consecutive) {
        memcpy(ctx->h_pin, h_sources, (size_t)n_sources * 4);
        NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_sources, ctx->h_pin, (size_t)n_sources * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    NEC_CUDA(ctx, cudaMemsetAsync(ctx->d_src_status, 0, ret_bytes, ctx->stream));
    // (the partial loads are zeroed by the CTAs that own them)
    nec::MidParams p{};
    p.n = n; p.n_pad = (n + 15) / 16 * 16; p.row_ptr = g->d_row_ptr; p.col_idx = g->d_col_idx; p.slab = g->d_slab; p.deg8 = g->d_deg8;
    p.sources = consecutive ? nullptr : ctx->d_sources; p.source_base = h_sources[0];
    p.n_sources = n_sources; p.include_endpoints = include_endpoints;
    p.q_base = d_q; p.q_stride = q_stride; p.order_base = d_order; p.order_stride = o_stride;
    p.pos4_base = d_pos4;
    p.bslot_base = d_bslot; p.bslot_stride = q_stride;
    p.source_status = d_status; p.counters = d_counters;
    if (mo) { p.measure_only = 1; p.out_ecc = mo->ecc; p.out_sumdist = mo->sumdist; p.out_reached = mo->reached; }
    if (const char *env = getenv("NEC_MID_CARVEOUT")) {          // A/B knob: preferred shared-memory carve-out in percent (what is left is L1)
        const int pct = atoi(env);
        cudaFuncSetAttribute(nec::mid_engine<8>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(nec::mid_engine<16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(nec::mid_engine<8, nec::kMidThreadsSmall>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(nec::mid_engine<16, nec::kMidThreadsSmall>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (small_ctas) {
        if (g->slab_w == 8) nec::mid_engine<8, nec::kMidThreadsSmall><<<slots, threads, smem, ctx->stream>>>(p);
        else nec::mid_engine<16, nec::kMidThreadsSmall><<<slots, threads, smem, ctx->stream>>>(p);
    } else {
        if (g->slab_w == 8) nec::mid_engine<8><<<slots, threads, smem, ctx->stream>>>(p);
        else nec::mid_engine<16><<<slots, threads, smem, ctx->stream>>>(p);
    }
    NEC_CUDA(ctx, cudaGetLastError());
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev_mid, ctx->stream));
    if (!mo) {
        // small graphs: a thread per vertex would be a handful of blocks, each thread walking all the slots
        if ((n + 255) / 256 < 148u && slots >= 64u)
            nec::reduce_slots_grouped_kernel<32><<<std::min<uint32_t>((n + 31) / 32, 148u * 2u), 1024, 0, ctx->stream>>>(d_bslot, q_stride, slots, n, d_out);
        else
            nec::reduce_slots_kernel<<<std::min<uint32_t>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(d_bslot, q_stride, slots, n, d_out);
        NEC_CUDA(ctx, cudaGetLastError());
    }
    NEC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NEC_CUDA(ctx, cudaMemcpyAsync(h_ret, ctx->d_src_status, ret_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    const auto t_host1 = std::chrono::steady_clock::now();
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (timing) {
        const auto t_host2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[nec] mid engine host path: issue %.1f us, wait %.1f us\n", std::chrono::duration<double, std::micro>(t_host1 - t_host0).count(),
                std::chrono::duration<double, std::micro>(t_host2 - t_host1).count());
    }
    unsigned long long counters[8];
    memcpy(counters, h_ret, sizeof counters);
    const uint8_t *status = h_ret + 64;
    float ms = 0.f;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->stats.kernel_ms = ms;
    NEC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev_mid));
    ctx->stats.engine_ms = ms;
    ctx->stats.kernel_launches = 2;
    ctx->stats.engine = 2;
    ctx->stats.team_size = 1;
    ctx->stats.cta_threads = (uint32_t)threads;
    ctx->stats.n_slots = slots;
    ctx->stats.workspace_bytes = ctx->mid_bytes;
    for (uint32_t k = 0; k < n_sources; ++k) {
        if (status[k] == nec::kMidTooDeep || status[k] == nec::kMidPredOverflow || status[k] == nec::kMidQueueOverflow) {
            fallback.push_back(h_sources[k]);
            if (fallback_index) fallback_index->push_back(k);
        }
        else if (status[k] != nec::kMidDone) return fail(ctx, NEC_EINTERNAL, "source index %u was not processed (status %u)", k, (unsigned)status[k]);
    }
    ctx->stats.reached_pairs = counters[0];
    ctx->stats.edge_pairs = counters[1];
    ctx->stats.vertex_visits = counters[2];
    ctx->stats.max_depth = (uint32_t)counters[3];
    const double tot = (double)(counters[4] + counters[5] + counters[6]);
    ctx->stats.frac_reset = tot > 0 ? (float)(counters[4] / tot) : 0.f;
    ctx->stats.frac_forward = tot > 0 ? (float)(counters[5] / tot) : 0.f;
    ctx->stats.frac_backward = tot > 0 ? (float)(counters[6] / tot) : 0.f;
    if (getenv("NEC_MID_VERBOSE")) fprintf(stderr, "[nec] mid engine: thread-0 time in levels of <= %d vertices: %.1f %% of CTA time\n", threads, 100.0 * counters[7] / tot);
    return NEC_OK;
}

// Which of the two engines suits a graph that FITS the mid engine is a property of its structure, not of its size: what the compact
// batched engine pays per source is proportional to the number of distinct BFS levels a vertex sits on within a batch (queue
// entries per vertex: ~4 on random and scale-free graphs, where it is 1.3-5 x faster than one source per CTA from n = 8192 up;
// 9 and more on small-world rings and lattices, where it is slower: scripts/exp_engine_choice.py, profiles/r2_engine_choice.log).
// So the first large call on such a graph runs a probe (the forward pass of vertex ids 0..255 as four 64-source batches, < 1 ms) and the finding is
// kept with the graph: a pure function of the graph, hence the same engine - and bitwise the same result - call after call.
constexpr uint32_t kProbeMinN = 8192;            // below: the mid engine's small CTAs win whatever the structure
constexpr uint32_t kProbeMinSources = 512;       // below: not worth a probe; one wave of the mid engine
constexpr float kProbeMaxEntries = 7.5f;         // queue entries per vertex of a 64-source batch; measured break-even: 7.0 -> 1.5 x for wide, 7.8 -> 1.0, 8.4 -> 0.97, 8.8 -> 0.77
int engine_hint(nec_ctx *ctx, const nec_graph *g, int *hint) {
    if (g->engine_hint == 0) {
        const uint32_t cnt = std::min<uint32_t>(g->n, 256u);
        std::vector<uint32_t> probe(cnt);
        for (uint32_t k = 0; k < cnt; ++k) probe[k] = k;
        int rc = grow(ctx, ctx->d_tmp, ctx->d_tmp_cap, g->n);
        if (rc != NEC_OK) return rc;
        {
            struct ProbeMode { nec_ctx *c; explicit ProbeMode(nec_ctx *x) : c(x) { c->probe_mode = true; } ~ProbeMode() { c->probe_mode = false; } } guard(ctx);
            rc = run_sources_wide(ctx, g, probe.data(), cnt, 1, ctx->d_tmp);
        }
        if (rc == NEC_ENOMEM) { g->engine_hint = 2; g->probe_entries = -1.f; }          // no room for a slot: the mid engine it is
        else if (rc != NEC_OK) return rc;
        else {
            const nec_stats &st = ctx->stats;
            g->probe_entries = st.requeued_batches ? 1e9f : (float)((double)st.vertex_visits / ((double)std::max<uint64_t>(st.n_batches, 1) * (double)std::max<uint32_t>(g->n, 1u)));
            g->engine_hint = g->probe_entries <= kProbeMaxEntries ? 4 : 2;
        }
        if (getenv("NEC_PROBE_VERBOSE")) fprintf(stderr, "[nec] probe: n=%u, %.2f queue entries per vertex in one batch -> %s engine\n", g->n, g->probe_entries, g->engine_hint == 4 ? "compact batched" : "mid");
        if (g->engine_hint == 2 && ctx->wide.base && ctx->wide.bytes > (1ull << 30)) {       // the probe's slots are of no further use (small ones stay for the next graph's probe)
            NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->wide.base);
            ctx->wide = WideArena{};
        }
    }
    *hint = g->engine_hint;
    return NEC_OK;
}
// does a call of n_sources sources on this (mid-capable) graph go to the compact batched engine instead?
int prefers_wide(nec_ctx *ctx, const nec_graph *g, uint64_t n_sources, bool *wide) {
    *wide = false;
    if (ctx->force_engine != 0 || g->n < kProbeMinN || n_sources < kProbeMinSources || (uint64_t)g->n * 256 + 256 >= 0xFFFFFFFFull) return NEC_OK;
    if (const char *env = getenv("NEC_WIDE")) { if (atoi(env) == 0) return NEC_OK; }
    if (const char *env = getenv("NEC_PROBE")) { if (atoi(env) == 0) return NEC_OK; }     // A/B knob: size threshold only
    int hint = 2;
    const int rc = engine_hint(ctx, g, &hint);
    if (rc != NEC_OK) return rc;
    *wide = hint == 4;
    return NEC_OK;
}

// Engine choice: graphs whose distance vector fits in shared memory take the one-source-per-CTA engine unless their structure
// suits the compact batched engine (engine_hint); larger ones take the compact batched engine; tiny ones, debug planes and
// whatever the others hand back take the source-minor batched engine.
int run_sources(nec_ctx *ctx, const nec_graph *g, const uint32_t *h_sources, uint32_t n_sources,
                int include_endpoints, double *d_out, bool debug, uint32_t *dbg_npred, double *dbg_bk) {
    if (const char *env = getenv("NEC_L2_FETCH")) {       // A/B knob: L2 fetch granularity hint (32 / 64 / 128 B)
        NEC_CUDA(ctx, cudaSetDevice(ctx->device));
        NEC_CUDA(ctx, cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(env)));
    }
    for (uint32_t k = 0; k < n_sources; ++k)
        if (h_sources[k] >= g->n) return fail(ctx, NEC_EINVAL, "source %u (index %u) out of range for n=%u", h_sources[k], k, g->n);
    const bool mid_ok = !debug && g->d_slab && g->n <= nec::kMidMaxN && nec::mid_smem_bytes(g->n, g->slab_w) <= 227u * 1024u - 4096u && n_sources > 0 && ctx->force_engine != 1 && ctx->force_engine != 3 &&
                        (g->n >= nec::kMidMinN || ctx->force_engine == 2) && g->n > 0;
    // large graphs (and whatever a test forces): the compact engine; tiny graphs, debug planes: the source-minor one
    bool wide_ok = !debug && !mid_ok && n_sources > 0 && g->n > 0 && (ctx->force_engine == 3 || (ctx->force_engine == 0 && g->n > nec::kMidMaxN)) &&
                   (uint64_t)g->n * 64 + 64 < 0xFFFFFFFFull;
    if (const char *env = getenv("NEC_WIDE")) { if (atoi(env) == 0 && ctx->force_engine != 3) wide_ok = false; }
    if (wide_ok) return run_sources_wide(ctx, g, h_sources, n_sources, include_endpoints, d_out);
    if (mid_ok) {
        bool wide = false;
        const int rcp = prefers_wide(ctx, g, n_sources, &wide);
        if (rcp != NEC_OK) return rcp;
        if (wide) return run_sources_wide(ctx, g, h_sources, n_sources, include_endpoints, d_out);
    }
    if (!mid_ok) return run_sources_batched(ctx, g, h_sources, n_sources, include_endpoints, d_out, debug, dbg_npred, dbg_bk);
    std::vector<uint32_t> fallback;
    int rc = run_sources_mid(ctx, g, h_sources, n_sources, include_endpoints, d_out, fallback);
    if (rc != NEC_OK || fallback.empty()) return rc;
    // the few sources the mid engine could not encode (depth > 126, > 62 predecessors, queue overflow)
    const nec_stats first = ctx->stats;
    if ((rc = grow(ctx, ctx->d_tmp, ctx->d_tmp_cap, g->n)) != NEC_OK) return rc;
    double *d_tmp = ctx->d_tmp;
    rc = run_sources_batched(ctx, g, fallback.data(), (uint32_t)fallback.size(), include_endpoints, d_tmp, false, nullptr, nullptr);
    if (rc != NEC_OK) return rc;
    nec::add_into_kernel<<<std::min<uint32_t>((g->n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(d_out, d_tmp, g->n);
    NEC_CUDA(ctx, cudaGetLastError());
    nec_stats &st = ctx->stats;                      // merge the two passes
    st.n_sources += first.n_sources - fallback.size();
    st.n_batches += first.n_batches;
    st.reached_pairs += first.reached_pairs; st.edge_pairs += first.edge_pairs; st.vertex_visits += first.vertex_visits;
    st.max_depth = std::max(st.max_depth, first.max_depth);
    st.kernel_launches += first.kernel_launches + 1;
    st.kernel_ms += first.kernel_ms; st.engine_ms += first.engine_ms;
    st.wide_batches += (uint32_t)fallback.size();
    st.engine = 1u | (2u << 4);
    st.workspace_bytes = std::max(st.workspace_bytes, first.workspace_bytes);
    return NEC_OK;
}

// ---- NCCL, resolved at run time -------------------------------------------------------------------------------
// Only a multi-device context needs it, and a host process may already carry its own copy (torch bundles one):
// dlopen by soname picks up whatever libnccl.so.2 the process has loaded, else the system library.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        auto sym = [&](const char *n) { void *p = dlsym(handle, n); if (!p) error = std::string("libnccl lacks ") + n; return p; };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) { handle = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;

#define NEC_NCCL(ctx, call)                                                                            \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess) return fail(ctx, NEC_ECUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

inline nec_ctx *device_ctx(nec_ctx *ctx, size_t d) { return d == 0 ? ctx : ctx->peers[d - 1]; }
inline const nec_graph *device_graph(const nec_graph *g, size_t d) { return d == 0 ? g : g->replicas[d - 1]; }
inline size_t n_devices_of(const nec_ctx *ctx) { return 1 + ctx->peers.size(); }

// Runs f(device index) on one host thread per device of a multi-device context and returns the first failure,
// its message copied into the primary context.
template <class F>
int on_all_devices(nec_ctx *ctx, F &&f) {
    const size_t nd = n_devices_of(ctx);
    std::vector<int> rcs(nd, NEC_OK);
    std::vector<std::thread> workers;
    workers.reserve(nd - 1);
    for (size_t d = 1; d < nd; ++d) workers.emplace_back([&, d] { rcs[d] = f(d); });
    rcs[0] = f(0);
    for (auto &w : workers) w.join();
    for (size_t d = 0; d < nd; ++d)
        if (rcs[d] != NEC_OK) {
            if (d) ctx->err = "device " + std::to_string(device_ctx(ctx, d)->device) + ": " + device_ctx(ctx, d)->err;
            cudaSetDevice(ctx->device);
            return rcs[d];
        }
    cudaSetDevice(ctx->device);
    return NEC_OK;
}

// Sources of a call, dealt to the devices in blocks of 64 (one wide batch), block b to device b mod D: per-source
// cost is about uniform inside a component, and neighbouring ids stay together (ring-like graphs).
std::vector<std::vector<uint32_t>> deal_sources(const uint32_t *sources, uint32_t n_sources, size_t nd) {
    std::vector<std::vector<uint32_t>> out(nd);
    for (auto &v : out) v.reserve(n_sources / nd + 64);
    for (uint32_t k = 0; k < n_sources; ++k) out[(k / 64) % nd].push_back(sources[k]);
    return out;
}

// Partial loads of all devices of a multi-device context -> the full vector in every device's d_out:
// ONE ncclAllReduce(sum, f64, n) over NVLink, issued for all devices from this thread inside a group.  Each
// device's reduce_slots kernel wrote its partial straight into d_out (the NCCL buffer), in place.
int all_reduce_loads(nec_ctx *ctx, uint32_t n) {
    const size_t nd = n_devices_of(ctx);
    NEC_NCCL(ctx, g_nccl.GroupStart());
    for (size_t d = 0; d < nd; ++d) {
        nec_ctx *c = device_ctx(ctx, d);
        ncclResult_t r = g_nccl.AllReduce(c->d_out, c->d_out, n, ncclDouble, ncclSum, ctx->comms[d], c->stream);
        if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(ctx, NEC_ECUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); }
    }
    NEC_NCCL(ctx, g_nccl.GroupEnd());
    return NEC_OK;
}

void merge_device_stats(nec_ctx *ctx) {
    nec_stats total = ctx->stats;
    for (nec_ctx *pc : ctx->peers) {
        const nec_stats &s = pc->stats;
        total.n_sources += s.n_sources; total.n_batches += s.n_batches; total.reached_pairs += s.reached_pairs;
        total.edge_pairs += s.edge_pairs; total.vertex_visits += s.vertex_visits;
        total.max_depth = std::max(total.max_depth, s.max_depth);
        total.kernel_launches += s.kernel_launches; total.wide_batches += s.wide_batches; total.requeued_batches += s.requeued_batches;
        total.n_slots += s.n_slots;
        total.kernel_ms = std::max(total.kernel_ms, s.kernel_ms); total.engine_ms = std::max(total.engine_ms, s.engine_ms);
        total.workspace_bytes += s.workspace_bytes;
        total.pull_levels += s.pull_levels; total.push_levels += s.push_levels;
    }
    ctx->stats = total;
}

int init_single(nec_ctx **out, int dev);

}  // namespace

extern "C" {

const char *nec_version(void) { return "net_ensembles_cuda 0.1.0 sm_100a"; }

const char *nec_last_error(const nec_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

// for host_bridge.cpp (same library): failures detected above the C ABI are reported through the context too
void nec_internal_set_error(nec_ctx *ctx, const char *msg) { if (ctx) ctx->err = msg ? msg : ""; }

int nec_init(nec_ctx **out, const int *devices, int n_devices) {
    if (!out) return fail(nullptr, NEC_EINVAL, "nec_init: out is NULL");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return fail(nullptr, NEC_EINVAL, "nec_init: n_devices must be 1..64 (got %d)", n_devices);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NEC_ENODEVICE, "no CUDA device available (%s); net_ensembles_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        if (devices) devs[i] = devices[i];
        else if (n_devices > 1) devs[i] = i;
        else if (cudaGetDevice(&devs[i]) != cudaSuccess) devs[i] = 0;
        if (devs[i] < 0 || devs[i] >= count) return fail(nullptr, NEC_EINVAL, "nec_init: device %d out of range (have %d)", devs[i], count);
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(nullptr, NEC_EINVAL, "nec_init: device %d listed twice", devs[i]);
    }
    nec_ctx *ctx = nullptr;
    int rc = init_single(&ctx, devs[0]);
    if (rc != NEC_OK) return rc;
    if (n_devices > 1) {
        // one full single-device context per further GPU, and one NCCL communicator per device (single process:
        // ncclCommInitAll), for the one all-reduce that combines the devices' partial load vectors
        if (!g_nccl.load()) { rc = fail(nullptr, NEC_ECUDA, "nec_init: %s", g_nccl.error.c_str()); nec_destroy(ctx); return rc; }
        for (int i = 1; i < n_devices; ++i) {
            nec_ctx *pc = nullptr;
            if ((rc = init_single(&pc, devs[i])) != NEC_OK) { nec_destroy(ctx); return rc; }
            ctx->peers.push_back(pc);
        }
        ctx->comms.assign(n_devices, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(ctx->comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) {
            rc = fail(nullptr, NEC_ECUDA, "nec_init: ncclCommInitAll over %d devices failed: %s", n_devices, g_nccl.GetErrorString(r));
            ctx->comms.clear();
            nec_destroy(ctx);
            return rc;
        }
        cudaSetDevice(devs[0]);
    }
    *out = ctx;
    return NEC_OK;
}

}  // extern "C"

namespace {
int init_single(nec_ctx **out, int dev) {
    cudaError_t e;
    nec_ctx *ctx = new (std::nothrow) nec_ctx();
    if (!ctx) return fail(nullptr, NEC_ENOMEM, "nec_init: out of host memory");
    ctx->device = dev;
    auto bail = [&](int code, const char *what, cudaError_t err) {
        int rc = fail(nullptr, code, "nec_init: %s: %s", what, cudaGetErrorString(err));
        nec_destroy(ctx);
        return rc;
    };
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return bail(NEC_ECUDA, "cudaSetDevice", e);
    if ((e = cudaGetDeviceProperties(&ctx->prop, dev)) != cudaSuccess) return bail(NEC_ECUDA, "cudaGetDeviceProperties", e);
    if (ctx->prop.major != 10) {
        int rc = fail(nullptr, NEC_ENODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", dev, ctx->prop.major, ctx->prop.minor);
        nec_destroy(ctx);
        return rc;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(NEC_ECUDA, "cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(NEC_ECUDA, "cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(NEC_ECUDA, "cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev_mid)) != cudaSuccess) return bail(NEC_ECUDA, "cudaEventCreate", e);
    if ((e = cudaMalloc((void **)&ctx->d_counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(NEC_ENOMEM, "cudaMalloc", e);
    {   // keep freed graph buffers in the device's default memory pool instead of returning them to the OS
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (engine_occupancy(ctx) != NEC_OK) {
        g_init_error = ctx->err;
        nec_destroy(ctx);
        return NEC_ECUDA;
    }
    *out = ctx;
    return NEC_OK;
}
}  // namespace

extern "C" {

void nec_destroy(nec_ctx *ctx) {
    if (!ctx) return;
    for (ncclComm_t c : ctx->comms)
        if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    ctx->comms.clear();
    for (nec_ctx *pc : ctx->peers) nec_destroy(pc);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->arena.base) cudaFree(ctx->arena.base);
    if (ctx->wide.base) cudaFree(ctx->wide.base);
    if (ctx->d_sources) cudaFree(ctx->d_sources);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_worklist) cudaFree(ctx->d_worklist);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_out) cudaFree(ctx->d_out);
    for (void *ptr : {(void *)ctx->d_keep, (void *)ctx->d_tmp, (void *)ctx->d_mecc, (void *)ctx->d_mreach, (void *)ctx->d_msum,
                      (void *)ctx->d_pecc, (void *)ctx->d_preach, (void *)ctx->d_psum, (void *)ctx->d_scatter})
        if (ptr) cudaFree(ptr);
    if (ctx->mid_base) cudaFree(ctx->mid_base);
    if (ctx->d_src_status) cudaFree(ctx->d_src_status);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->small.release();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_sample) cudaFree(ctx->d_sample);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int nec_set_stream(nec_ctx *ctx, void *cuda_stream) {
    if (!ctx) return NEC_EINVAL;
    if (cuda_stream && !ctx->peers.empty()) return fail(ctx, NEC_EINVAL, "nec_set_stream: a multi-device context works on its own streams (one per device)");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NEC_OK;
}

int nec_set_engine(nec_ctx *ctx, int engine) {
    if (!ctx || engine < 0 || engine > 3) return NEC_EINVAL;
    ctx->force_engine = engine;
    for (nec_ctx *pc : ctx->peers) pc->force_engine = engine;
    return NEC_OK;
}

int nec_set_workspace_limit(nec_ctx *ctx, uint64_t bytes) {
    if (!ctx) return NEC_EINVAL;
    ctx->ws_limit = bytes;
    for (nec_ctx *pc : ctx->peers) pc->ws_limit = bytes;
    return NEC_OK;
}

// Shared tail of the two upload entry points: the 32-bit CSR sits validated in the context's pinned staging
// area (row pointers, then neighbour ids); allocate the device graph, copy, and build the mid engine's slab
// rows and degree bytes ON THE DEVICE from the CSR that has just arrived.
static int finish_upload_one(nec_ctx *ctx, uint32_t n, uint64_t nnz, const uint32_t *h_rp32, const uint32_t *h_col,
                             uint32_t max_deg, size_t over8, nec_graph **out);

// ... on every device of the context (each copies from the same pinned staging area over its own link)
static int finish_upload(nec_ctx *ctx, uint32_t n, uint64_t nnz, const uint32_t *h_rp32, const uint32_t *h_col,
                         uint32_t max_deg, size_t over8, nec_graph **out) {
    if (ctx->peers.empty()) return finish_upload_one(ctx, n, nnz, h_rp32, h_col, max_deg, over8, out);
    std::vector<nec_graph *> made(n_devices_of(ctx), nullptr);
    int rc = on_all_devices(ctx, [&](size_t d) { return finish_upload_one(device_ctx(ctx, d), n, nnz, h_rp32, h_col, max_deg, over8, &made[d]); });
    if (rc != NEC_OK) {
        for (size_t d = 0; d < made.size(); ++d) nec_graph_free(device_ctx(ctx, d), made[d]);
        return rc;
    }
    made[0]->replicas.assign(made.begin() + 1, made.end());
    *out = made[0];
    return NEC_OK;
}

static int finish_upload_one(nec_ctx *ctx, uint32_t n, uint64_t nnz, const uint32_t *h_rp32, const uint32_t *h_col,
                             uint32_t max_deg, size_t over8, nec_graph **out) {
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    nec_graph *g = new (std::nothrow) nec_graph();
    if (!g) return fail(ctx, NEC_ENOMEM, "out of host memory");
    g->n = n; g->nnz = nnz; g->max_degree = max_deg;
    // stream-ordered pool allocations: a Markov chain uploads a fresh graph per sample, and plain
    // cudaMalloc/cudaFree pairs (device-wide synchronisation, up to 0.5 s observed) would dominate
    cudaError_t e = cudaMallocAsync((void **)&g->d_row_ptr, ((size_t)n + 1) * sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&g->d_col_idx, std::max<size_t>(nnz, 1) * sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->d_row_ptr, h_rp32, ((size_t)n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(g->d_col_idx, h_col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n > 0 && n <= nec::kMidMaxN) {
        // slab rows: [degree, first W-1 neighbours]; W = 8 unless more than 10 % of the rows would overflow
        g->slab_w = over8 * 10 > n ? 16 : 8;
        const size_t slab_words = (size_t)n * g->slab_w;
        e = cudaMallocAsync((void **)&g->d_slab, slab_words * sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&g->d_deg8, (size_t)n + 2, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->d_deg8 + n, 0, 2, ctx->stream);
        if (e == cudaSuccess) {
            const unsigned blocks = (unsigned)std::min<size_t>((slab_words + 255) / 256, 148u * 16u);
            nec::build_slab_kernel<<<blocks, 256, 0, ctx->stream>>>(n, g->slab_w, g->d_row_ptr, g->d_col_idx, reinterpret_cast<uint32_t *>(g->d_slab), g->d_deg8);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      // the staging area is reused by the next upload
    if (e != cudaSuccess) {
        nec_graph_free(ctx, g);
        return fail(ctx, e == cudaErrorMemoryAllocation ? NEC_ENOMEM : NEC_ECUDA, "graph upload: %s", cudaGetErrorString(e));
    }
    *out = g;
    return NEC_OK;
}

// Pinned staging of the 32-bit CSR: [row_ptr (n+1) | pad to 256 B | col_idx (nnz)].
static int stage_reserve(nec_ctx *ctx, uint32_t n, uint64_t nnz, uint32_t *&h_rp32, uint32_t *&h_col) {
    const size_t rp_bytes = align_up(((size_t)n + 1) * 4, 256), need = rp_bytes + std::max<uint64_t>(nnz, 1) * 4;
    if (need > ctx->h_stage_cap) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_sample) cudaFree(ctx->d_sample);
        ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
        const size_t cap = need + need / 8;
        cudaError_t e = cudaHostAlloc(&ctx->h_stage, cap, cudaHostAllocPortable);   // every device of a multi-device context copies from it
        if (e != cudaSuccess) return fail(ctx, NEC_ENOMEM, "cudaHostAlloc(%zu B CSR staging) failed: %s", cap, cudaGetErrorString(e));
        ctx->h_stage_cap = cap;
    }
    h_rp32 = reinterpret_cast<uint32_t *>(ctx->h_stage);
    h_col = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ctx->h_stage) + rp_bytes);
    return NEC_OK;
}

int nec_graph_upload(nec_ctx *ctx, uint32_t n, uint64_t nnz, const uint64_t *row_ptr, const uint32_t *col_idx, nec_graph **out) try {
    if (!ctx) return NEC_EINVAL;
    if (!out) return fail(ctx, NEC_EINVAL, "nec_graph_upload: out is NULL");
    *out = nullptr;
    if (!row_ptr || (nnz && !col_idx)) return fail(ctx, NEC_EINVAL, "nec_graph_upload: NULL CSR pointer");
    if (nnz >= 0xFFFFFFFFull) return fail(ctx, NEC_EINVAL, "nec_graph_upload: nnz=%llu does not fit 32-bit row pointers", (unsigned long long)nnz);
    if (n > 0x7FFFFFFFu / 64u * 16) return fail(ctx, NEC_EINVAL, "nec_graph_upload: n=%u too large", n);
    if ((uint64_t)n * 64u + 64u > 0xFFFFFFFFull) return fail(ctx, NEC_EINVAL, "nec_graph_upload: n=%u: the level queues (64 entries per vertex) need 32-bit positions", n);
    if (row_ptr[0] != 0 || row_ptr[n] != nnz) return fail(ctx, NEC_EINVAL, "nec_graph_upload: row_ptr[0] must be 0 and row_ptr[n] == nnz");
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t *rp32 = nullptr, *col32 = nullptr;
    int rc = stage_reserve(ctx, n, nnz, rp32, col32);
    if (rc != NEC_OK) return rc;
    uint32_t max_deg = 0;
    size_t over8 = 0;
    // validation, the 64 -> 32 bit row pointers and the copy into pinned memory in ONE pass per array, on all
    // host cores for large graphs (tens of ms at N = 2^22)
    long long bad_row = (long long)n, bad_entry = (long long)nnz;
#pragma omp parallel for schedule(static) reduction(max : max_deg) reduction(min : bad_row) reduction(+ : over8) if (n > 100000)
    for (long long v = 0; v < (long long)n; ++v) {
        if (row_ptr[v + 1] < row_ptr[v]) bad_row = std::min(bad_row, v);
        else {
            const uint64_t d = row_ptr[v + 1] - row_ptr[v];
            max_deg = std::max<uint32_t>(max_deg, (uint32_t)std::min<uint64_t>(d, 0xFFFFFFFFull));
            over8 += d > 7;
        }
        rp32[v] = (uint32_t)row_ptr[v];
    }
    if (bad_row < (long long)n) return fail(ctx, NEC_EINVAL, "nec_graph_upload: row_ptr not monotone at %u", (uint32_t)bad_row);
    rp32[n] = (uint32_t)nnz;
#pragma omp parallel for schedule(static) reduction(min : bad_entry) if (nnz > 400000)
    for (long long e = 0; e < (long long)nnz; ++e) {
        const uint32_t x = col_idx[e];
        if (x >= n) bad_entry = std::min(bad_entry, e);
        col32[e] = x;
    }
    if (bad_entry < (long long)nnz)
        return fail(ctx, NEC_EINVAL, "nec_graph_upload: col_idx[%llu]=%u out of range (n=%u)", (unsigned long long)bad_entry, col_idx[bad_entry], n);
    return finish_upload(ctx, n, nnz, rp32, col32, max_deg, over8, out);
} NEC_CATCH_ALL(ctx)

// Export fast path: the crate keeps a vertex's neighbours in one contiguous slice (`AdjList::edges()`:
// &[usize] for NodeContainer, src/graph.rs:145-150; &[SwEdge] for SwContainer, src/sw_graph.rs:164-170), so
// the caller hands over n (pointer, length) pairs plus the element layout and the library flattens them -
// narrowing to u32, validating and writing straight into its pinned staging area on all host cores - instead
// of the caller building a CSR in pageable memory that is then copied again.  Adjacency ORDER is preserved.
int nec_graph_upload_adjacency(nec_ctx *ctx, uint32_t n, const void *const *rows, const uint64_t *degrees, uint32_t elem_stride,
                               uint32_t to_offset, uint32_t to_width, nec_graph **out) try {
    if (!ctx) return NEC_EINVAL;
    if (!out) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: out is NULL");
    *out = nullptr;
    if (n && (!rows || !degrees)) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: NULL argument");
    if ((to_width != 4 && to_width != 8) || elem_stride < to_offset + to_width)
        return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: element layout stride=%u offset=%u width=%u is not a 4- or 8-byte id inside the element", elem_stride, to_offset, to_width);
    if ((uint64_t)n * 64u + 64u > 0xFFFFFFFFull) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: n=%u too large", n);
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t nnz = 0;
    uint32_t max_deg = 0;
    size_t over8 = 0;
    for (uint32_t v = 0; v < n; ++v) {
        if (degrees[v] && !rows[v]) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: row %u is NULL with %llu entries", v, (unsigned long long)degrees[v]);
        nnz += degrees[v];
        if (nnz >= 0xFFFFFFFFull) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: more than 2^32 - 2 adjacency entries");
        max_deg = std::max<uint32_t>(max_deg, (uint32_t)degrees[v]);
        over8 += degrees[v] > 7;
    }
    uint32_t *rp32 = nullptr, *col32 = nullptr;
    int rc = stage_reserve(ctx, n, nnz, rp32, col32);
    if (rc != NEC_OK) return rc;
    {
        uint64_t at = 0;
        for (uint32_t v = 0; v < n; ++v) { rp32[v] = (uint32_t)at; at += degrees[v]; }
        rp32[n] = (uint32_t)at;
    }
    long long bad_row = (long long)n;
#pragma omp parallel for schedule(dynamic, 4096) reduction(min : bad_row) if (nnz > 400000)
    for (long long v = 0; v < (long long)n; ++v) {
        const uint8_t *src = static_cast<const uint8_t *>(rows[v]) + to_offset;
        uint32_t *dst = col32 + rp32[v];
        const uint64_t d = degrees[v];
        bool ok = true;
        if (to_width == 8)
            for (uint64_t k = 0; k < d; ++k) { uint64_t x; std::memcpy(&x, src + k * elem_stride, 8); ok = ok && x < n; dst[k] = (uint32_t)x; }
        else
            for (uint64_t k = 0; k < d; ++k) { uint32_t x; std::memcpy(&x, src + k * elem_stride, 4); ok = ok && x < n; dst[k] = x; }
        if (!ok) bad_row = std::min(bad_row, v);
    }
    if (bad_row < (long long)n) return fail(ctx, NEC_EINVAL, "nec_graph_upload_adjacency: a neighbour id of vertex %u is out of range (n=%u)", (uint32_t)bad_row, n);
    return finish_upload(ctx, n, nnz, rp32, col32, max_deg, over8, out);
} NEC_CATCH_ALL(ctx)

// ErEnsembleC::randomize on one device; the generator state is read from / written to 4 x u64 (state lo, hi,
// increment lo, hi).  Deterministic, so a multi-device context simply runs it on every device.
static int erc_randomize_one(nec_ctx *ctx, uint32_t n, double prob, const uint64_t rng[4], nec_graph **out) {
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    nec_graph *g = new (std::nothrow) nec_graph();
    if (!g) return fail(ctx, NEC_ENOMEM, "out of host memory");
    g->n = n;
    auto bail = [&](int rc) { nec_graph_free(ctx, g); return rc; };
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_state = up((size_t)n * 16), b_u32 = up((size_t)n * 4), b_tot = 256, b_step = up(33 * sizeof(nec::LcgStep));
    const size_t need = b_state + 3 * b_u32 + b_tot + b_step;
    int rc = grow(ctx, ctx->d_sample, ctx->d_sample_cap, need);
    if (rc != NEC_OK) return bail(rc);
    nec::ErcParams p{};
    p.n = n; p.prob = prob;
    p.state_lo = rng[0]; p.state_hi = rng[1]; p.inc_lo = rng[2]; p.inc_hi = rng[3];
    uint8_t *d = ctx->d_sample;
    p.row_state = (unsigned long long *)d; d += b_state;
    p.lower_deg = (uint32_t *)d; d += b_u32;
    p.upper_deg = (uint32_t *)d; d += b_u32;
    p.cursor = (uint32_t *)d; d += b_u32;
    p.totals = (unsigned long long *)d; d += b_tot;
    nec::LcgStep *d_steps = (nec::LcgStep *)d;
    p.lane_step = d_steps;
    nec::LcgStep steps[33];
    const nec::u128 inc = nec::u128_of(rng[2], rng[3]);
    for (int l = 0; l < 33; ++l) {
        nec::u128 m, a;
        nec::lcg_jump(l < 32 ? (unsigned long long)l + 1 : 32ull, inc, m, a);
        steps[l] = nec::LcgStep{(unsigned long long)m, (unsigned long long)(m >> 64), (unsigned long long)a, (unsigned long long)(a >> 64)};
    }
    cudaError_t e = cudaMallocAsync((void **)&g->d_row_ptr, ((size_t)n + 1) * sizeof(uint32_t), ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, NEC_ENOMEM, "nec_erc_randomize: %s", cudaGetErrorString(e)));
    p.row_ptr = g->d_row_ptr;
    NEC_CUDA(ctx, cudaMemcpyAsync(d_steps, steps, sizeof steps, cudaMemcpyHostToDevice, ctx->stream));
    NEC_CUDA(ctx, cudaMemsetAsync(p.totals, 0, 3 * sizeof(unsigned long long), ctx->stream));
    const unsigned row_blocks = std::min<unsigned>((n + 7) / 8, 148u * 8u);       // 8 warps per CTA, one warp per row
    nec::erc_row_states_kernel<<<std::min<unsigned>((n + 255) / 256, 148u * 4u), 256, 0, ctx->stream>>>(p);
    nec::erc_rows_kernel<false><<<row_blocks, 256, 0, ctx->stream>>>(p);
    nec::erc_scan_kernel<<<1, 1024, 0, ctx->stream>>>(p);
    unsigned long long totals[3] = {0, 0, 0};
    e = cudaMemcpyAsync(totals, p.totals, sizeof totals, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);          // nnz decides the size of col_idx (pageable read: waited for)
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return bail(fail(ctx, NEC_ECUDA, "nec_erc_randomize: %s", cudaGetErrorString(e)));
    if (totals[0] >= 0xFFFFFFFFull) return bail(fail(ctx, NEC_EINVAL, "nec_erc_randomize: %llu adjacency entries do not fit 32-bit row pointers", totals[0]));
    g->nnz = totals[0]; g->max_degree = (uint32_t)totals[2];
    e = cudaMallocAsync((void **)&g->d_col_idx, std::max<size_t>(g->nnz, 1) * sizeof(uint32_t), ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, NEC_ENOMEM, "nec_erc_randomize: %s", cudaGetErrorString(e)));
    p.col_idx = g->d_col_idx;
    nec::erc_rows_kernel<true><<<row_blocks, 256, 0, ctx->stream>>>(p);
    nec::erc_sort_lower_kernel<<<std::min<unsigned>((n + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(p);
    if (n <= nec::kMidMaxN) {
        g->slab_w = totals[1] * 10 > n ? 16 : 8;
        const size_t slab_words = (size_t)n * g->slab_w;
        e = cudaMallocAsync((void **)&g->d_slab, slab_words * sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&g->d_deg8, (size_t)n + 2, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->d_deg8 + n, 0, 2, ctx->stream);
        if (e != cudaSuccess) return bail(fail(ctx, NEC_ENOMEM, "nec_erc_randomize: %s", cudaGetErrorString(e)));
        nec::build_slab_kernel<<<(unsigned)std::min<size_t>((slab_words + 255) / 256, 148u * 16u), 256, 0, ctx->stream>>>(
            n, g->slab_w, g->d_row_ptr, g->d_col_idx, reinterpret_cast<uint32_t *>(g->d_slab), g->d_deg8);
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return bail(fail(ctx, NEC_ECUDA, "nec_erc_randomize: %s", cudaGetErrorString(e)));
    *out = g;
    return NEC_OK;
}

int nec_erc_randomize(nec_ctx *ctx, uint32_t n, double prob, const uint64_t rng_in[4], uint64_t rng_out[4], nec_graph **out) try {
    if (!ctx) return NEC_EINVAL;
    if (!out || !rng_in) return fail(ctx, NEC_EINVAL, "nec_erc_randomize: NULL argument");
    *out = nullptr;
    if (n < 2) return fail(ctx, NEC_EINVAL, "nec_erc_randomize: n must be at least 2");
    if ((uint64_t)n * 64u + 64u > 0xFFFFFFFFull) return fail(ctx, NEC_EINVAL, "nec_erc_randomize: n=%u too large", n);
    if (!(rng_in[2] & 1)) return fail(ctx, NEC_EINVAL, "nec_erc_randomize: a Pcg64 increment is odd");
    int rc;
    if (ctx->peers.empty()) rc = erc_randomize_one(ctx, n, prob, rng_in, out);
    else {
        std::vector<nec_graph *> made(n_devices_of(ctx), nullptr);
        rc = on_all_devices(ctx, [&](size_t d) { return erc_randomize_one(device_ctx(ctx, d), n, prob, rng_in, &made[d]); });
        if (rc != NEC_OK) for (size_t d = 0; d < made.size(); ++d) nec_graph_free(device_ctx(ctx, d), made[d]);
        else { made[0]->replicas.assign(made.begin() + 1, made.end()); *out = made[0]; }
    }
    if (rc != NEC_OK) return rc;
    if (rng_out) {      // the generator after its n (n - 1) / 2 draws
        nec::u128 m, a;
        const nec::u128 inc = nec::u128_of(rng_in[2], rng_in[3]);
        nec::lcg_jump((unsigned long long)n * (n - 1) / 2, inc, m, a);
        const nec::u128 s = m * nec::u128_of(rng_in[0], rng_in[1]) + a;
        rng_out[0] = (uint64_t)s; rng_out[1] = (uint64_t)(s >> 64); rng_out[2] = rng_in[2]; rng_out[3] = rng_in[3];
    }
    return NEC_OK;
} NEC_CATCH_ALL(ctx)

int nec_graph_dims(const nec_graph *g, uint32_t *n, uint64_t *nnz) {
    if (!g) return NEC_EINVAL;
    if (n) *n = g->n;
    if (nnz) *nnz = g->nnz;
    return NEC_OK;
}

int nec_graph_download(nec_ctx *ctx, const nec_graph *g, uint64_t *row_ptr, uint32_t *col_idx) try {
    if (!ctx) return NEC_EINVAL;
    if (!g || !row_ptr || (g->nnz && !col_idx)) return fail(ctx, NEC_EINVAL, "nec_graph_download: NULL argument");
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<uint32_t> rp32((size_t)g->n + 1);
    NEC_CUDA(ctx, cudaMemcpyAsync(rp32.data(), g->d_row_ptr, rp32.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (g->nnz) NEC_CUDA(ctx, cudaMemcpyAsync(col_idx, g->d_col_idx, g->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t v = 0; v <= g->n; ++v) row_ptr[v] = rp32[v];
    return NEC_OK;
} NEC_CATCH_ALL(ctx)

void nec_graph_free(nec_ctx *ctx, nec_graph *g) {
    if (!g) return;
    for (size_t i = 0; i < g->replicas.size(); ++i)
        nec_graph_free(ctx && i < ctx->peers.size() ? ctx->peers[i] : nullptr, g->replicas[i]);
    g->replicas.clear();
    if (ctx) {
        cudaSetDevice(ctx->device);
        if (g->d_row_ptr) cudaFreeAsync(g->d_row_ptr, ctx->stream);
        if (g->d_col_idx) cudaFreeAsync(g->d_col_idx, ctx->stream);
        if (g->d_slab) cudaFreeAsync(g->d_slab, ctx->stream);
        if (g->d_deg8) cudaFreeAsync(g->d_deg8, ctx->stream);
    }
    delete g;
}

// Multi-device call: the sources are dealt to the devices, every device runs its share on its own host thread,
// stream and arena into its d_out, one all-reduce combines them; afterwards every device's d_out holds the result.
static int run_sources_all_devices(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources, int include_endpoints) {
    const size_t nd = n_devices_of(ctx);
    if (g->replicas.size() + 1 != nd) return fail(ctx, NEC_EINVAL, "graph was not uploaded through this multi-device context");
    for (uint32_t k = 0; k < n_sources; ++k)
        if (sources[k] >= g->n) return fail(ctx, NEC_EINVAL, "source %u (index %u) out of range for n=%u", sources[k], k, g->n);
    const auto share = deal_sources(sources, n_sources, nd);
    int rc = on_all_devices(ctx, [&](size_t d) {
        nec_ctx *c = device_ctx(ctx, d);
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, NEC_ECUDA, "cudaSetDevice(%d) failed", c->device);
        int r = grow(c, c->d_out, c->d_out_cap, g->n);
        if (r != NEC_OK) return r;
        return run_sources(c, device_graph(g, d), share[d].data(), (uint32_t)share[d].size(), include_endpoints, c->d_out, false, nullptr, nullptr);
    });
    if (rc != NEC_OK) return rc;
    merge_device_stats(ctx);
    if ((rc = all_reduce_loads(ctx, g->n)) != NEC_OK) return rc;
    ctx->stats.kernel_launches += (uint32_t)nd;      // the collective's kernels, one per device
    return NEC_OK;
}

static int sync_all_devices(nec_ctx *ctx) {
    for (size_t d = n_devices_of(ctx); d-- > 0;) {
        nec_ctx *c = device_ctx(ctx, d);
        NEC_CUDA(ctx, cudaSetDevice(c->device));
        NEC_CUDA(ctx, cudaStreamSynchronize(c->stream));
    }
    return NEC_OK;
}

int nec_vertex_load_sources_device(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources,
                                   int include_endpoints, double *d_out) try {
    if (!ctx) return NEC_EINVAL;
    if (!g || (g->n && !d_out) || (n_sources && !sources)) return fail(ctx, NEC_EINVAL, "nec_vertex_load_sources_device: NULL argument");
    if (ctx->peers.empty()) return run_sources(ctx, g, sources, n_sources, include_endpoints ? 1 : 0, d_out, false, nullptr, nullptr);
    if (g->n == 0) { ctx->stats = nec_stats{}; return NEC_OK; }
    int rc = run_sources_all_devices(ctx, g, sources, n_sources, include_endpoints ? 1 : 0);
    if (rc != NEC_OK) return rc;
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));        // d_out lives on the context's first device
    NEC_CUDA(ctx, cudaMemcpyAsync(d_out, ctx->d_out, (size_t)g->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return sync_all_devices(ctx);
} NEC_CATCH_ALL(ctx)

int nec_vertex_load_sources(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources,
                            int include_endpoints, double *out) try {
    if (!ctx) return NEC_EINVAL;
    if (!g || (g->n && !out) || (n_sources && !sources)) return fail(ctx, NEC_EINVAL, "nec_vertex_load_sources: NULL argument");
    if (g->n == 0) { ctx->stats = nec_stats{}; return NEC_OK; }
    if (!ctx->peers.empty()) {
        int rc = run_sources_all_devices(ctx, g, sources, n_sources, include_endpoints ? 1 : 0);
        if (rc != NEC_OK) return rc;
        NEC_CUDA(ctx, cudaSetDevice(ctx->device));
        NEC_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out, (size_t)g->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));   // D2H once
        return sync_all_devices(ctx);
    }
    int rc = grow(ctx, ctx->d_out, ctx->d_out_cap, g->n);
    if (rc != NEC_OK) return rc;
    rc = run_sources(ctx, g, sources, n_sources, include_endpoints ? 1 : 0, ctx->d_out, false, nullptr, nullptr);
    if (rc != NEC_OK) return rc;
    NEC_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out, (size_t)g->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NEC_OK;
} NEC_CATCH_ALL(ctx)

int nec_vertex_load(nec_ctx *ctx, const nec_graph *g, int include_endpoints, double *out) try {
    if (!ctx) return NEC_EINVAL;
    if (!g) return fail(ctx, NEC_EINVAL, "nec_vertex_load: graph is NULL");
    std::vector<uint32_t> all(g->n);
    for (uint32_t v = 0; v < g->n; ++v) all[v] = v;   // generic_graph.rs:1321 — every vertex is a source
    return nec_vertex_load_sources(ctx, g, all.data(), g->n, include_endpoints, out);
} NEC_CATCH_ALL(ctx)

int nec_bfs_debug(nec_ctx *ctx, const nec_graph *g, uint32_t source, uint32_t *dist, uint32_t *npred, double *sigma, double *bk) try {
    if (!ctx) return NEC_EINVAL;
    if (!g) return fail(ctx, NEC_EINVAL, "nec_bfs_debug: graph is NULL");
    const uint32_t n = g->n;
    if (source >= n) return fail(ctx, NEC_EINVAL, "nec_bfs_debug: source %u out of range (n=%u)", source, n);
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    // Runs the PRODUCTION engine (debug instantiation: identical code plus two plane stores)
    // on a one-source batch and reads lane 0 of the slot's planes.
    uint32_t *d_np_plane = nullptr, *d_u32 = nullptr;
    double *d_bk_plane = nullptr, *d_sig_plane = nullptr, *d_f64 = nullptr, *d_load = nullptr;
    const size_t plane = (size_t)n * nec::kLanes;
    auto cleanup = [&]() { cudaFree(d_np_plane); cudaFree(d_bk_plane); cudaFree(d_sig_plane); cudaFree(d_u32); cudaFree(d_f64); cudaFree(d_load); ctx->dbg_sigma = nullptr; };
    cudaError_t e = cudaMalloc((void **)&d_np_plane, plane * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_bk_plane, plane * 8);
    if (e == cudaSuccess && sigma) e = cudaMalloc((void **)&d_sig_plane, plane * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_u32, (size_t)n * 4 * 2);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_f64, (size_t)n * 8 * 2);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_load, (size_t)n * 8);
    if (e != cudaSuccess) { cleanup(); return fail(ctx, NEC_ENOMEM, "nec_bfs_debug: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(d_np_plane, 0, plane * 4, ctx->stream);
    if (d_sig_plane) cudaMemsetAsync(d_sig_plane, 0, plane * 8, ctx->stream);      // unreached vertices: 0 paths
    nec::fill_f64_kernel<<<std::min<size_t>((plane + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(d_bk_plane, plane, 1.0);
    ctx->dbg_sigma = d_sig_plane;
    int rc = run_sources(ctx, g, &source, 1, 1, d_load, true, d_np_plane, d_bk_plane);
    if (rc != NEC_OK) { cleanup(); return rc; }
    const Arena &a = ctx->arena;
    const uint32_t blocks = std::min<uint32_t>((n + 255) / 256, 148u * 8u);
    uint32_t *d_dist32 = d_u32, *d_np = d_u32 + n;
    if (a.dist_bytes == 1) nec::extract_lane_kernel<uint8_t, uint32_t><<<blocks, 256, 0, ctx->stream>>>((const uint8_t *)a.dist, n, 0, d_dist32);
    else nec::extract_lane_kernel<uint32_t, uint32_t><<<blocks, 256, 0, ctx->stream>>>((const uint32_t *)a.dist, n, 0, d_dist32);
    if (a.dist_bytes == 1) nec::widen_inf_kernel<<<blocks, 256, 0, ctx->stream>>>(d_dist32, n);
    nec::extract_lane_kernel<uint32_t, uint32_t><<<blocks, 256, 0, ctx->stream>>>(d_np_plane, n, 0, d_np);
    nec::extract_lane_kernel<double, double><<<blocks, 256, 0, ctx->stream>>>(d_bk_plane, n, 0, d_f64);
    if (d_sig_plane) nec::extract_lane_kernel<double, double><<<blocks, 256, 0, ctx->stream>>>(d_sig_plane, n, 0, d_f64 + n);
    if (dist) cudaMemcpyAsync(dist, d_dist32, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (npred) cudaMemcpyAsync(npred, d_np, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (bk) cudaMemcpyAsync(bk, d_f64, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (sigma) cudaMemcpyAsync(sigma, d_f64 + n, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(ctx, NEC_ECUDA, "nec_bfs_debug: %s", cudaGetErrorString(e));
    return NEC_OK;
} NEC_CATCH_ALL(ctx)

// Distance measures of `sources` on ONE device (the context's own): forward pass of the engine the graph size selects.
static int distance_measures_one_device(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources,
                                        uint32_t *ecc, uint64_t *sum_dist, uint32_t *reached) {
    ctx->stats = nec_stats{};
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = grow(ctx, ctx->d_mecc, ctx->d_mecc_cap, n_sources)) != NEC_OK) return rc;
    if ((rc = grow(ctx, ctx->d_msum, ctx->d_msum_cap, n_sources)) != NEC_OK) return rc;
    if ((rc = grow(ctx, ctx->d_mreach, ctx->d_mreach_cap, n_sources)) != NEC_OK) return rc;
    MeasureOut mo{ctx->d_mecc, ctx->d_msum, ctx->d_mreach};
    for (uint32_t k = 0; k < n_sources; ++k)
        if (sources[k] >= g->n) return fail(ctx, NEC_EINVAL, "source %u (index %u) out of range for n=%u", sources[k], k, g->n);
    const bool mid_ok = g->d_slab && g->n <= nec::kMidMaxN && ctx->force_engine != 1 && (g->n >= nec::kMidMinN || ctx->force_engine == 2) &&
                        nec::mid_smem_bytes(g->n, g->slab_w) <= 227u * 1024u - 4096u;
    if (mid_ok) {
        std::vector<uint32_t> fb, fb_index;
        rc = run_sources_mid(ctx, g, sources, n_sources, 1, nullptr, fb, &mo, &fb_index);
        if (rc == NEC_OK && !fb.empty()) {       // sources deeper than 126 levels: batched engine, scattered back on the stream
            const nec_stats first = ctx->stats;
            const size_t k = fb.size();
            if ((rc = grow(ctx, ctx->d_pecc, ctx->d_pecc_cap, k)) != NEC_OK) return rc;
            if ((rc = grow(ctx, ctx->d_psum, ctx->d_psum_cap, k)) != NEC_OK) return rc;
            if ((rc = grow(ctx, ctx->d_preach, ctx->d_preach_cap, k)) != NEC_OK) return rc;
            if ((rc = grow(ctx, ctx->d_scatter, ctx->d_scatter_cap, k)) != NEC_OK) return rc;
            MeasureOut part{ctx->d_pecc, ctx->d_psum, ctx->d_preach};
            rc = run_sources_batched(ctx, g, fb.data(), (uint32_t)k, 1, nullptr, false, nullptr, nullptr, &part);
            if (rc != NEC_OK) return rc;
            NEC_CUDA(ctx, cudaMemcpyAsync(ctx->d_scatter, fb_index.data(), k * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            nec::scatter_measures_kernel<<<(unsigned)std::min<size_t>((k + 255) / 256, 148u * 8u), 256, 0, ctx->stream>>>(
                (uint32_t)k, ctx->d_scatter, part.ecc, part.sumdist, part.reached, mo.ecc, mo.sumdist, mo.reached);
            NEC_CUDA(ctx, cudaGetLastError());
            NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // fb_index is a pageable host vector
            ctx->stats.kernel_ms += first.kernel_ms; ctx->stats.engine_ms += first.engine_ms;
            ctx->stats.edge_pairs += first.edge_pairs; ctx->stats.reached_pairs += first.reached_pairs;
            ctx->stats.kernel_launches += first.kernel_launches + 1;
            ctx->stats.n_sources = n_sources; ctx->stats.engine = 1u | (2u << 4);
        }
    } else {
        rc = run_sources_batched(ctx, g, sources, n_sources, 1, nullptr, false, nullptr, nullptr, &mo);
    }
    if (rc != NEC_OK) return rc;
    if (ecc) NEC_CUDA(ctx, cudaMemcpyAsync(ecc, mo.ecc, (size_t)n_sources * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (sum_dist) NEC_CUDA(ctx, cudaMemcpyAsync(sum_dist, mo.sumdist, (size_t)n_sources * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (reached) NEC_CUDA(ctx, cudaMemcpyAsync(reached, mo.reached, (size_t)n_sources * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NEC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NEC_OK;
}

int nec_distance_measures(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources,
                          uint32_t *ecc, uint64_t *sum_dist, uint32_t *reached) try {
    if (!ctx) return NEC_EINVAL;
    if (!g || (n_sources && !sources)) return fail(ctx, NEC_EINVAL, "nec_distance_measures: NULL argument");
    ctx->stats = nec_stats{};
    if (n_sources == 0 || g->n == 0) return NEC_OK;
    if (!ctx->peers.empty()) {
        // per-source outputs: contiguous shares of the source list, one per device, nothing to combine
        const size_t nd = n_devices_of(ctx);
        if (g->replicas.size() + 1 != nd) return fail(ctx, NEC_EINVAL, "graph was not uploaded through this multi-device context");
        int rcm = on_all_devices(ctx, [&](size_t d) {
            const uint32_t k0 = (uint32_t)((uint64_t)n_sources * d / nd), k1 = (uint32_t)((uint64_t)n_sources * (d + 1) / nd);
            nec_ctx *c = device_ctx(ctx, d);
            nec_graph shallow = *device_graph(g, d);          // a single-device view of this device's replica
            shallow.replicas.clear();
            if (k1 == k0) { c->stats = nec_stats{}; return (int)NEC_OK; }
            return distance_measures_one_device(c, &shallow, sources + k0, k1 - k0, ecc ? ecc + k0 : nullptr,
                                                sum_dist ? sum_dist + k0 : nullptr, reached ? reached + k0 : nullptr);
        });
        if (rcm == NEC_OK) merge_device_stats(ctx);
        return rcm;
    }
    return distance_measures_one_device(ctx, g, sources, n_sources, ecc, sum_dist, reached);
} NEC_CATCH_ALL(ctx)

int nec_vertex_load_batch(nec_ctx *ctx, uint32_t n_graphs, const uint32_t *n_per_graph, const uint64_t *row_ptr_offsets,
                          const uint32_t *row_ptr_concat, const uint64_t *col_offsets, const uint32_t *col_idx_concat,
                          int include_endpoints, double *out_concat) try {
    if (!ctx) return NEC_EINVAL;
    if (n_graphs && (!n_per_graph || !row_ptr_offsets || !row_ptr_concat || !col_offsets || !out_concat))
        return fail(ctx, NEC_EINVAL, "nec_vertex_load_batch: NULL argument");
    ctx->stats = nec_stats{};
    if (!ctx->peers.empty() && n_graphs >= 2 * n_devices_of(ctx)) {
        // the batch shards by graph (contiguous ranges, no collective: the load vectors are just concatenated)
        const size_t nd = n_devices_of(ctx);
        std::vector<uint64_t> out_off(n_graphs + 1, 0);
        for (uint32_t i = 0; i < n_graphs; ++i) out_off[i + 1] = out_off[i] + n_per_graph[i];
        int rcm = on_all_devices(ctx, [&](size_t d) {
            const uint32_t g0 = (uint32_t)((uint64_t)n_graphs * d / nd), g1 = (uint32_t)((uint64_t)n_graphs * (d + 1) / nd);
            nec_ctx *c = device_ctx(ctx, d);
            c->stats = nec_stats{};
            if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, NEC_ECUDA, "cudaSetDevice(%d) failed", c->device);
            return nec::small_batch_run(c->small, c->stream, &c->err, &c->stats, g1 - g0, n_per_graph + g0, row_ptr_offsets + g0, row_ptr_concat,
                                        col_offsets + g0, col_idx_concat, include_endpoints ? 1 : 0, out_concat + out_off[g0]);
        });
        if (rcm == NEC_OK) merge_device_stats(ctx);
        return rcm;
    }
    cudaSetDevice(ctx->device);
    return nec::small_batch_run(ctx->small, ctx->stream, &ctx->err, &ctx->stats, n_graphs, n_per_graph,
                                row_ptr_offsets, row_ptr_concat, col_offsets, col_idx_concat, include_endpoints ? 1 : 0, out_concat);
} NEC_CATCH_ALL(ctx)

int nec_chains_create(nec_ctx *ctx, uint32_t n_graphs, const uint32_t *n_per_graph, const uint64_t *row_ptr_offsets,
                      const uint32_t *row_ptr_concat, const uint64_t *col_offsets, const uint32_t *col_idx_concat,
                      uint32_t slack, nec_chains **out) try {
    if (!ctx) return NEC_EINVAL;
    if (!out) return fail(ctx, NEC_EINVAL, "nec_chains_create: out is NULL");
    *out = nullptr;
    if (!n_graphs || !n_per_graph || !row_ptr_offsets || !row_ptr_concat || !col_offsets)
        return fail(ctx, NEC_EINVAL, "nec_chains_create: NULL argument or empty set");
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    return nec::chains_create(ctx->small, ctx->stream, &ctx->err, n_graphs, n_per_graph, row_ptr_offsets, row_ptr_concat, col_offsets,
                              col_idx_concat, slack, out);
} NEC_CATCH_ALL(ctx)

int nec_chains_step_and_load(nec_ctx *ctx, nec_chains *chains, const uint32_t *op_offsets, const uint64_t *ops, int include_endpoints,
                             double *out_concat) try {
    if (!ctx) return NEC_EINVAL;
    if (!chains) return fail(ctx, NEC_EINVAL, "nec_chains_step_and_load: chains is NULL");
    ctx->stats = nec_stats{};
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    return nec::chains_step_and_load(ctx->small, ctx->stream, &ctx->err, &ctx->stats, chains, op_offsets, ops, include_endpoints ? 1 : 0, out_concat);
} NEC_CATCH_ALL(ctx)

const double *nec_chains_result(const nec_chains *chains) { return chains ? chains->h_out : nullptr; }

int nec_chains_download(nec_ctx *ctx, nec_chains *chains, uint32_t *row_ptr_concat, uint32_t *col_idx_concat, uint64_t col_capacity, uint64_t *nnz) try {
    if (!ctx) return NEC_EINVAL;
    if (!chains || !row_ptr_concat || (col_capacity && !col_idx_concat)) return fail(ctx, NEC_EINVAL, "nec_chains_download: NULL argument");
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    return nec::chains_download(ctx->stream, &ctx->err, chains, row_ptr_concat, col_idx_concat, col_capacity, nnz);
} NEC_CATCH_ALL(ctx)

void nec_chains_free(nec_ctx *ctx, nec_chains *chains) {
    if (!chains) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    nec::chains_release(chains);
}

int nec_wave_size(nec_ctx *ctx, const nec_graph *g, uint64_t n_sources_total, uint32_t *sources_per_wave) try {
    if (!ctx) return NEC_EINVAL;
    if (!g || !sources_per_wave) return fail(ctx, NEC_EINVAL, "nec_wave_size: NULL argument");
    *sources_per_wave = 0;
    if (g->n == 0 || n_sources_total == 0) return NEC_OK;
    NEC_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t ns = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, n_sources_total / n_devices_of(ctx)), 0xFFFFFFFFull);   // per device
    const bool mid_ok = g->d_slab && g->n <= nec::kMidMaxN && nec::mid_smem_bytes(g->n, g->slab_w) <= 227u * 1024u - 4096u &&
                        ctx->force_engine != 1 && ctx->force_engine != 3 && (g->n >= nec::kMidMinN || ctx->force_engine == 2);
    bool mid_prefers_wide = false;
    if (mid_ok) {
        const int rcp = prefers_wide(ctx, g, ns, &mid_prefers_wide);
        if (rcp != NEC_OK) return rcp;
    }
    if (mid_ok && !mid_prefers_wide) {       // one source per CTA, two CTAs per SM
        uint32_t small_n = 16384;
        if (const char *env = getenv("NEC_MID_SMALL_N")) small_n = (uint32_t)atol(env);
        const uint64_t per_sm = g->n <= small_n ? 8ull : 2ull;      // CTAs (= sources in flight) per SM, see run_sources_mid
        *sources_per_wave = (uint32_t)std::min<uint64_t>(n_sources_total, std::min<uint64_t>(ns, per_sm * (uint32_t)ctx->prop.multiProcessorCount) * n_devices_of(ctx));
        return NEC_OK;
    }
    bool wide_ok = (ctx->force_engine == 3 || (ctx->force_engine == 0 && g->n > nec::kMidMaxN)) && (uint64_t)g->n * 64 + 64 < 0xFFFFFFFFull;
    if (const char *env = getenv("NEC_WIDE")) { if (atoi(env) == 0 && ctx->force_engine != 3) wide_ok = false; }
    if (wide_ok || mid_prefers_wide) {
        WidePlan wp{};
        const int rcw = plan_wide(ctx, g->n, ns, wp);
        if (rcw != NEC_OK) return rcw;
        *sources_per_wave = (uint32_t)std::min<uint64_t>(n_sources_total, (uint64_t)wp.slots * wp.kw * n_devices_of(ctx));
        return NEC_OK;
    }
    BatchedPlan plan{};
    const int rc = plan_batched(ctx, g->n, ns, false, false, plan);
    if (rc != NEC_OK) return rc;
    *sources_per_wave = (uint32_t)std::min<uint64_t>(n_sources_total, (uint64_t)plan.slots * plan.ksrc * n_devices_of(ctx));
    return NEC_OK;
} NEC_CATCH_ALL(ctx)

int nec_get_stats(const nec_ctx *ctx, nec_stats *out) {
    if (!ctx || !out) return NEC_EINVAL;
    *out = ctx->stats;
    return NEC_OK;
}

}  // extern "C"
