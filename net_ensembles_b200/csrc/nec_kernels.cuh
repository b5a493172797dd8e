// CUDA kernels of the vertex_load hot path (sm_100a).
//
// What is computed (reference: src/generic_graph/generic_graph.rs:1310-1385): for every
// source s a BFS gives d_s(.); with P_s(w) = {u in N(w): d_s(u) = d_s(w)-1} the reverse
// sweep yields  bk_s(w) = 1 + sum_{x: w in P_s(x)} bk_s(x)/|P_s(x)|  and the result is
// b[v] = sum_{s != v, s reaches v} (bk_s(v) - [!include_endpoints]).
//
// How it is laid out for the GPU — "batched multi-source, source-minor":
//   * B = 32 or 64 sources form a BATCH (template parameter kSrc; the host picks 64 when the call has
//     at least one such batch per SM and the workspace holds one 64-wide slot per SM).  Source k of
//     the batch is bit k of every mask and byte / element k of every row; lane k of a warp owns
//     source k (and k + 32 at B = 64).
//   * one CTA owns one batch SLOT (its arena of state) and walks the batch level by level
//     with __syncthreads only; CTAs are persistent and take batches round-robin (static
//     batch -> slot mapping, so every floating-point sum is reproducible run to run).
//   * per-slot state: one mask record per vertex {seen, next[even], next[odd], -} (u32 or u64
//     words; seen = sources that have reached v, next = sources that reach v on the level being
//     built, double-buffered by level parity, self-cleaning), dist[v][B] (u8; a 64 B row is one
//     DRAM burst; u32 rows for graphs deeper than 253 levels), q[v][B] (f64), the level queues
//     (vertex ids and their source masks) and a private partial load vector.
//   * FORWARD pass, bit-parallel.  Per level: (1) MARK - every queue entry's mask moves from its
//     next[] word into the queue and is ORed into seen (plain read-modify-write: one owner per
//     vertex and level), so that (2) EXPAND needs no atomic on seen: one LANE PER ADJACENCY ENTRY,
//     a warp takes a tile of 32 queue entries, writes their distance rows (lane k stores byte k,
//     one store instruction per row and 32 sources), flattens their adjacency lists (warp scan +
//     per-lane owner search) and every lane handles (u -> x) entries for all B sources at once,
//     four per step, each stage a batch of independent memory operations:
//     new = mask(u) & ~(seen[x] | next[x]) (a snapshot), atomicOr(next[x], new) merges racing
//     lanes exactly, and the lane that turns next[x] non-zero appends x to the next level's queue
//     (warp-aggregated shared-memory tail).
//   * BACKWARD pass (pull, atomic-free, deterministic), one LANE PER SOURCE: levels deepest
//     -> 1 over the same queues, again in tiles of 32 entries with flattened adjacency entries
//     consumed in chunks of 32 through a two-deep pipeline: chunk s+2's neighbour ids are being
//     loaded, chunk s+1's 32 distance rows are in flight to shared memory (cp.async, 16 B pieces),
//     chunk s is consumed.  For entry (w, mask) and neighbour x, lane k reads
//     byte k of x's staged row; lanes with dist == level+1 pull q[x][k], lanes count
//     dist == level-1 as predecessors; at the end of w's row bk = 1 + sum, q[w][k] = bk / npred
//     and the lane-sum of bk goes to the slot's private b[] (exactly one writer per vertex per
//     level => no atomics).  Predecessor lists are never materialised; sigma never enters the
//     value.  kMode 2 runs the forward pass only and reports per-source eccentricity, sum of
//     distances and reached count (nec_distance_measures).
//   * CTA shape follows residency: 2 x 512 threads per SM when slots are plentiful, 1 x 1024 when memory
//     allows about one slot per SM, and - when it allows fewer slots than SMs - a thread-block CLUSTER of
//     512-thread CTAs per slot (kTeam): the team splits each level's tiles, shares the queue tail and the
//     failure flag through distributed shared memory and meets at cluster barriers.
//   * slots' private b[] are summed in slot order by reduce_slots_kernel.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#ifndef NEC_Q_PREFETCH
#define NEC_Q_PREFETCH 0      // backward pass: 0 no prefetch of a chunk's q values, 1 into L2, 2 into L1 (A/B: both lose 19-25 %)
#endif
#ifndef NEC_LEAF_MARKS
#define NEC_LEAF_MARKS 0      // 1: leaf marks in the backward pass (A/B: 4-10 % slower although they halve the q loads)
#endif

namespace nec {

constexpr int kLanes = 32;            // lanes per warp = sources per narrow batch (debug planes, small-graph kernel)
// CTA shape: 512 threads x 2 CTAs per SM when the workspace budget allows two slots per SM, else
// 1024 threads x 1 CTA per SM (large graphs are slot-limited by memory; wider CTAs keep the SMs busy)
constexpr int kUnroll = 8;            // neighbour rows in flight per warp step (backward)
constexpr int kFwdUnroll = 4;         // flattened adjacency entries per lane and step (forward, push)
constexpr int kPullUnroll = 4;        // neighbours' frontier masks in flight per thread (forward, pull)
constexpr uint32_t kPullHubDeg = 64;  // longer adjacency lists are gathered by a whole warp in the pull scan

enum BatchStatus : uint8_t { kPending = 0, kDone = 1, kTooDeep = 2, kQueueOverflow = 3 };

struct EngineParams {
    // graph
    uint32_t n;
    const uint32_t *__restrict__ row_ptr;
    const uint32_t *__restrict__ col_idx;
    // work
    const uint32_t *__restrict__ sources;     // n_sources vertex ids
    uint32_t n_sources;
    const uint32_t *__restrict__ work_list;   // batch ids to run (nullptr: 0..n_work-1)
    uint32_t n_work;
    int include_endpoints;
    // forward pass direction per level: 0 always push (frontier -> neighbours, atomics on next[]), 1 cost model,
    // 2 always pull (every not-yet-complete vertex gathers its neighbours' frontier masks; no atomics)
    uint32_t pull_mode;
    uint32_t avg_deg;                         // ceil(nnz / n), >= 1 (cost model of the direction choice)
    // per-slot arenas (slot = blockIdx.x)
    uint8_t *dist_base;   size_t dist_stride;   // bytes per slot
    double *q_base;       size_t q_stride;      // doubles per slot
    uint8_t *mask_base;   size_t mask_stride;   // bytes per slot (n records of 4 masks: seen, next[even], next[odd], -; u32 or u64 each)
    double *bslot_base;   size_t bslot_stride;  // doubles per slot (n)
    uint32_t *queue_v_base;                     // level queues: vertex ids ...
    uint8_t *queue_m_base;                      // ... and their source masks (u32 or u64 per entry)
    size_t queue_stride;                        // entries per slot (pitch)
    size_t queue_cap;                           // usable entries per slot
    uint32_t *lvl_base;   size_t lvl_stride;    // u32 per slot (max_levels + 2)
    uint32_t max_levels;                        // deepest level index the slot can record
    // outputs
    uint8_t *batch_status;                      // per batch id
    unsigned long long *counters;               // [0] reached pairs [1] edge pairs [2] visits [3] max depth [4..6] CTA cycles in reset/forward/backward [7] pulled << 32 | pushed level expansions
    // optional per-(vertex,lane) debug planes of slot 0 (nec_bfs_debug)
    uint32_t *dbg_npred;
    double *dbg_bk;
    double *dbg_sigma;                          // f64 shortest-path counts (not used by the load: the crate splits equally)
    // distance-measure outputs, indexed like `sources` (kMode 2)
    uint32_t *out_ecc;
    unsigned long long *out_sumdist;
    uint32_t *out_reached;
};

template <typename DistT> struct DistTraits;
template <> struct DistTraits<uint8_t>  { static constexpr uint32_t kInf = 0xFFu;        static constexpr uint32_t kMaxLevel = 126u; };   // bytes >= 0x80: 'finished' marks of the backward pass
template <> struct DistTraits<uint32_t> { static constexpr uint32_t kInf = 0xFFFFFFFFu;  static constexpr uint32_t kMaxLevel = 0xFFFFFFF0u; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Predicated global atomic OR with return value: one ATOMG instruction under a predicate instead of
// a branch, so several of them can be in flight per thread.  Returns 0xFFFFFFFF when not executed.
__device__ __forceinline__ uint32_t atom_or_if(bool pred, uint32_t *addr, uint32_t val) {
    uint32_t old = 0xFFFFFFFFu;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.global.or.b32 %0, [%1], %2;\n\t}"
                 : "+r"(old) : "l"(addr), "r"(val), "r"((uint32_t)pred) : "memory");
    return old;
}

// Inclusive warp scan of per-lane counts.
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Flattened item t (0 <= t < incl[31]) belongs to the first lane whose inclusive prefix
// exceeds t; lanes with zero items are skipped automatically.  Branch-free 5-step search.
__device__ __forceinline__ int tile_owner(uint32_t incl, uint32_t t) {
    int lo = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, lo + step - 1);
        if (v <= t) lo += step;
    }
    return lo;
}

// Batch width: 32 sources (one per lane) or 64 (two per lane: sources `lane` and `lane + 32`).  The
// forward pass is bit-parallel, so its cost per adjacency entry does not depend on the width; the
// backward pass reads one distance row per adjacency entry whatever the width (a 64 B row is one full
// DRAM transaction where a 32 B row wastes half of one).  Wide batches double the per-slot state.
template <int kSrc> struct SrcTraits;
template <> struct SrcTraits<32> {
    using Mask = uint32_t;
    using Stage = uint2;                                  // (neighbour id, owner's source mask)
    static __device__ __forceinline__ Stage stage(uint32_t x, Mask m) { return make_uint2(x, m); }
    static __device__ __forceinline__ uint32_t half(const Stage &e, int) { return e.y; }
    static __device__ __forceinline__ bool any(const Stage &e) { return e.y != 0; }
    static __device__ __forceinline__ uint32_t popc(Mask m) { return (uint32_t)__popc(m); }
};
template <> struct SrcTraits<64> {
    using Mask = unsigned long long;
    using Stage = uint4;                                  // (neighbour id, mask low, mask high, -)
    static __device__ __forceinline__ Stage stage(uint32_t x, Mask m) { return make_uint4(x, (uint32_t)m, (uint32_t)(m >> 32), 0u); }
    static __device__ __forceinline__ uint32_t half(const Stage &e, int h) { return h ? e.z : e.y; }
    static __device__ __forceinline__ bool any(const Stage &e) { return (e.y | e.z) != 0; }
    static __device__ __forceinline__ uint32_t popc(Mask m) { return (uint32_t)__popcll(m); }
};

__device__ __forceinline__ unsigned long long atom_or_if(bool pred, unsigned long long *addr, unsigned long long val) {
    unsigned long long old = ~0ull;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.global.or.b64 %0, [%1], %2;\n\t}"
                 : "+l"(old) : "l"(addr), "l"(val), "r"((uint32_t)pred) : "memory");
    return old;
}

// The engine.  One CTA = one slot; loops over its batches.
// kMode 0: vertex_load; 1: vertex_load + per-(vertex,lane) debug planes (nec_bfs_debug);
// 2: distance measures only — the forward pass alone, per source eccentricity / sum of distances /
//    reached count (diameter, closeness_centrality: reference generic_graph.rs:1116-1144, :1387-1403).
template <typename DistT, int kThreads, int kSrc>
constexpr size_t engine_smem_bytes() {
    return (size_t)(kThreads / 32) * (2 * 32 * sizeof(typename SrcTraits<kSrc>::Stage) + (sizeof(DistT) == 1 ? 2 * 32 * kSrc : 0));
}

// kTeam: the CTAs of a thread-block CLUSTER work on one slot together (cluster size from the launch).  Used when
// memory holds fewer slots than there are SMs (config 5: 96 slots on 148 SMs): three 512-thread CTAs per slot
// put 1536 threads on each batch.  The team shares the queue tail and the failure flag through distributed
// shared memory (rank 0's copies) and meets at cluster barriers instead of __syncthreads; everything else is
// per warp and unchanged, and since every (vertex, source) value still has exactly one writer and the same
// summation order, results are bit-identical to the single-CTA kernel.
template <typename DistT, int kMode, int kEngineThreads, int kSrc, bool kTeam = false>
__global__ void __launch_bounds__(kEngineThreads, kEngineThreads == 512 ? 2 : 1)
vertex_load_engine(const EngineParams p) {
    static_assert(!kTeam || kMode == 0, "cluster teams run the vertex_load mode only");
    using ST = SrcTraits<kSrc>;
    using Mask = typename ST::Mask;
    using Stage = typename ST::Stage;
    constexpr int kH = kSrc / 32;                        // sources per lane
    constexpr int kUnrollB = kUnroll / kH;               // neighbour rows per backward step (kUnroll loads in flight per lane)
    constexpr bool kDebug = kMode == 1;
    constexpr bool kMeasure = kMode == 2;
    constexpr uint32_t kInf = DistTraits<DistT>::kInf;
    constexpr int kWarps = kEngineThreads / 32;
    static_assert(!kDebug || kSrc == 32, "debug planes are 32 sources wide");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t team_size = 1, team_rank = 0;               // CTAs per slot, this CTA's rank among them
    if constexpr (kTeam) {
        team_size = cooperative_groups::this_cluster().num_blocks();
        team_rank = cooperative_groups::this_cluster().block_rank();
    }
    const uint32_t slot = blockIdx.x / team_size, n_slots = gridDim.x / team_size;
    const uint32_t tid_t = team_rank * kEngineThreads + threadIdx.x;      // thread and warp index within the team
    const uint32_t threads_t = team_size * kEngineThreads;
    const uint32_t warp_t = team_rank * kWarps + warp, warps_t = team_size * kWarps;
    auto team_sync = [&]() {
        if constexpr (kTeam) cooperative_groups::this_cluster().sync();
        else __syncthreads();
    };
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t n = p.n;

    DistT *__restrict__ dist = reinterpret_cast<DistT *>(p.dist_base + (size_t)slot * p.dist_stride);
    double *__restrict__ q = p.q_base + (size_t)slot * p.q_stride;
    // per-vertex mask record {seen, next[even level], next[odd level], -}: one 16 B / 32 B sector, so the
    // atomics that follow the `seen` test of a neighbour hit the sector that test just fetched
    Mask *rec = reinterpret_cast<Mask *>(p.mask_base + (size_t)slot * p.mask_stride);
    double *__restrict__ bslot = p.bslot_base + (size_t)slot * p.bslot_stride;
    uint32_t *__restrict__ queue_v = p.queue_v_base + (size_t)slot * p.queue_stride;
    Mask *__restrict__ queue_m = reinterpret_cast<Mask *>(p.queue_m_base + (size_t)slot * p.queue_stride * sizeof(Mask));
    uint32_t *__restrict__ lvl = p.lvl_base + (size_t)slot * p.lvl_stride;
    const uint32_t qcap = (uint32_t)p.queue_cap;
    const uint32_t level_limit = p.max_levels < DistTraits<DistT>::kMaxLevel ? p.max_levels : DistTraits<DistT>::kMaxLevel;

    __shared__ uint32_t s_tail;
    __shared__ uint32_t s_fail;
    __shared__ uint32_t s_open;                          // pull scans count the vertices that still miss a source
    // a team's queue tail and failure flag are rank 0's copies, reached through distributed shared memory;
    // a single CTA uses its own shared variables directly (shared-space atomics, no generic addressing)
    uint32_t *tail_p = nullptr, *fail_p = nullptr, *open_p = nullptr;
    if constexpr (kTeam) {
        tail_p = cooperative_groups::this_cluster().map_shared_rank(&s_tail, 0);
        fail_p = cooperative_groups::this_cluster().map_shared_rank(&s_fail, 0);
        open_p = cooperative_groups::this_cluster().map_shared_rank(&s_open, 0);
    }
    auto open_add = [&](uint32_t cnt) {
        if constexpr (kTeam) atomicAdd(open_p, cnt);
        else atomicAdd(&s_open, cnt);
    };
    auto open_read = [&]() -> uint32_t {
        if constexpr (kTeam) return *(volatile uint32_t *)open_p;
        else return s_open;
    };
    auto tail_add = [&](uint32_t cnt) -> uint32_t {
        if constexpr (kTeam) return atomicAdd(tail_p, cnt);
        else return atomicAdd(&s_tail, cnt);
    };
    auto tail_read = [&]() -> uint32_t {
        if constexpr (kTeam) return *(volatile uint32_t *)tail_p;
        else return s_tail;
    };
    auto fail_set = [&]() {
        if constexpr (kTeam) *(volatile uint32_t *)fail_p = 1;
        else s_fail = 1;
    };
    auto fail_read = [&]() -> bool {
        if constexpr (kTeam) return *(volatile uint32_t *)fail_p != 0;
        else return s_fail != 0;
    };
    // LEAF MARKS (byte distance rows, vertex_load mode): when the backward pass finishes (w, source k) it overwrites
    // w's distance byte with 0x80 | parity(level) << 6 | info.  info = npred for a BFS leaf (no successors: bk = 1
    // exactly, so q = 1 / npred and neither a q store nor - for the vertices that pull from it - a random 8-byte q
    // load is needed: the puller has the byte in its staged row already and takes 1 / npred from a 64-entry table);
    // info = 0 for everything else (q is read as before).  About half of all BFS-DAG edges of an expander end in a
    // leaf.  Readers on level l see successors as marks of parity(l + 1), predecessors as plain bytes l - 1 and
    // same-level vertices as plain l or marks of parity(l), whichever wins the race: all three stay distinct.
    constexpr bool kLeafMarks = NEC_LEAF_MARKS && sizeof(DistT) == 1 && kMode == 0;
    constexpr uint32_t kMaxLeafPred = 62;
    __shared__ double s_rcp[64];
    __shared__ unsigned char s_klist[kWarps][64];      // backward: the live sources of the entry a warp is consuming
    if (kLeafMarks && threadIdx.x < 64) s_rcp[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __shared__ unsigned long long s_msum[kSrc];          // kMode 2: per source of the batch
    __shared__ uint32_t s_mreach[kSrc], s_mecc[kSrc];
    __shared__ unsigned long long s_cnt[3];
    constexpr bool kStageRows = sizeof(DistT) == 1;      // byte distance rows are staged through shared memory
    extern __shared__ __align__(16) unsigned char engine_smem[];
    // backward: (neighbour, owner's source mask) per flattened entry; staged distance rows
    Stage (*s_stage)[2][32] = reinterpret_cast<Stage (*)[2][32]>(engine_smem);
    unsigned char (*s_rows)[2][32][kSrc] = reinterpret_cast<unsigned char (*)[2][32][kSrc]>(engine_smem + (size_t)kWarps * 2 * 32 * sizeof(Stage));

    uint32_t my_depth = 0;
    long long cyc_reset = 0, cyc_fwd = 0, cyc_bwd = 0;      // thread 0 only
    unsigned long long dir_levels = 0;                      // thread 0 only: pulled levels << 32 | pushed levels
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;

    for (uint32_t work = slot; work < p.n_work; work += n_slots) {
        const uint32_t batch = p.work_list ? p.work_list[work] : work;
        uint32_t b_reached = 0, b_edges = 0, b_visits = 0;   // per lane, per batch; counted only if the batch completes

        long long t_phase = clock64();
        // ---- reset: only the mask records (whole records: full-sector stores; the next[] words are zero
        //      already).  Stale dist/q are never used: the backward pass selects source k of a row only if
        //      k reached that vertex in this batch. -----
        {
            uint4 *r16 = reinterpret_cast<uint4 *>(rec);
            const size_t n16 = (size_t)n * 4 * sizeof(Mask) / 16;
            for (size_t i = tid_t; i < n16; i += threads_t) __stcg(&r16[i], make_uint4(0u, 0u, 0u, 0u));
        }
        if (kDebug) {
            const size_t n16 = ((size_t)n * kSrc * sizeof(DistT)) / 16;
            uint4 *d16 = reinterpret_cast<uint4 *>(dist);
            const uint4 inf4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            for (size_t i = tid_t; i < n16; i += threads_t) d16[i] = inf4;
        }
        if (threadIdx.x == 0) { s_tail = 0; s_fail = 0; s_open = 0; }
        if (kMeasure && threadIdx.x < kSrc) { s_msum[threadIdx.x] = 0; s_mreach[threadIdx.x] = 0; s_mecc[threadIdx.x] = 0; }
        team_sync();

        // ---- seed level 0: bit k <-> source k of the batch ------------------------------
        if (warp_t == 0) {
#pragma unroll
            for (int h = 0; h < kH; ++h) {
                const uint32_t k = h * 32 + lane;
                const uint32_t sidx = batch * kSrc + k;
                if (sidx < p.n_sources) {
                    const uint32_t s = p.sources[sidx];
                    dist[(size_t)s * kSrc + k] = (DistT)0;
                    const Mask old = atomicOr(&rec[(size_t)s * 4 + 1], (Mask)1 << k);
                    if (old == 0) queue_v[tail_add(1u)] = s;
                }
            }
        }

        team_sync();
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_reset += t - t_phase; t_phase = t; }
        // ---- forward: level-synchronous, bit-parallel; direction chosen per level -------
        // PUSH (narrow levels): one lane per adjacency entry of the frontier, atomicOr into next[] of the neighbour.
        // PULL (wide levels): one thread per vertex that still misses a source ORs the frontier masks of its
        // neighbours - no atomics, one writer per vertex, the vertex records / row pointers / adjacency lists are
        // read as streams and only the neighbours' 8-byte frontier masks are random reads.  Both produce the same
        // level sets, masks and (up to order within a level, which no sum depends on) queues.
        uint32_t level = 0, qbeg = 0, qend = 0;
        bool too_deep = false;
        uint32_t n_open = n;                         // vertices that miss a source of the batch (upper bound until a pull scan counts them)
        uint32_t defer_beg = 0, defer_end = 0;       // queue range of a pulled level: its next[] words are cleared one level later
        bool prev_pull = false;
        Mask valid_src = ~(Mask)0;                   // sources this batch really has
        {
            const uint32_t first = batch * kSrc;
            const uint32_t have = p.n_sources - first < (uint32_t)kSrc ? p.n_sources - first : (uint32_t)kSrc;
            if (have < (uint32_t)kSrc) valid_src = ((Mask)1 << have) - 1;
        }
        for (;;) {
            team_sync();                             // previous level fully expanded
            const uint32_t tail_raw = tail_read();
            const uint32_t open_raw = open_read();
            const uint32_t tail_now = tail_raw < qcap ? tail_raw : qcap;
            team_sync();                             // every warp sampled the tail before it moves again
            qbeg = qend;
            qend = tail_now;
            if (prev_pull) n_open = open_raw;
            if (tid_t == 0) lvl[level] = qbeg;
            if (qend == qbeg) break;                 // level `level` is empty: done
            if (level >= level_limit) { too_deep = true; break; }
            Mask *next_cur = rec + 1 + (level & 1);          // word of the record, indexed [4 * vertex]
            Mask *next_nxt = rec + 2 - (level & 1);
            const DistT cur_d = (DistT)level;
            // direction: a pushed row costs a random record read plus (mostly) an atomic, a pulled row a random
            // 8-byte read; the pull scan also streams all records once
            const uint32_t entries = qend - qbeg;
            const bool pull = p.pull_mode == 2 ||
                              (p.pull_mode == 1 && 2ull * entries * p.avg_deg > (9ull * n) / 16 + (unsigned long long)n_open * p.avg_deg);

            // mark: the level's vertices become `seen` for their sources BEFORE anything is expanded, so the
            // expansion needs no atomic on `seen` (a plain test suffices: bits of this and earlier levels are
            // final, and a source discovered twice on the level being built is caught by the atomic on next[]).
            // The entry's mask moves to the queue (the backward pass wants it there) and next[] cleans itself -
            // one level later if this level is pulled (the scan reads the masks from the records).
            for (uint32_t idx = qbeg + tid_t; idx < qend; idx += threads_t) {
                const uint32_t u = queue_v[idx];
                const Mask m = __ldcg(&next_cur[(size_t)u * 4]);
                const Mask sn = __ldcg(&rec[(size_t)u * 4]);
                __stcg(&rec[(size_t)u * 4], sn | m);
                if (!pull) __stcg(&next_cur[(size_t)u * 4], (Mask)0);   // this parity is reused at level+2
                queue_m[idx] = m;
            }
            for (uint32_t idx = defer_beg + tid_t; idx < defer_end; idx += threads_t)
                __stcg(&next_nxt[(size_t)queue_v[idx] * 4], (Mask)0);   // the pulled level-1's masks: same parity as level+1
            defer_beg = pull ? qbeg : 0u;
            defer_end = pull ? qend : 0u;
            if (pull && threadIdx.x == 0) s_open = 0;
            team_sync();

            for (uint32_t tile = qbeg + warp_t * 32; tile < qend; tile += warps_t * 32) {
                const uint32_t idx = tile + lane;
                Mask m = 0;
                uint32_t rs = 0, deg = 0, u = 0;
                if (idx < qend) {
                    u = queue_v[idx];
                    m = queue_m[idx];
                    rs = __ldg(p.row_ptr + u);
                    deg = __ldg(p.row_ptr + u + 1) - rs;
                    const uint32_t pc = ST::popc(m);
                    b_reached += pc;
                    b_edges += pc * deg;
                    b_visits += 1;
                }
                // distance rows of the dequeued entries: the mask is final now, so each row's bytes are
                // written by ONE store instruction per 32 sources (lane k = sources k, k+32) instead of
                // bit by bit at discovery
                if (kMeasure) {
                    // source k has popc(ballot(bit k)) vertices of this tile on the current level
                    uint32_t mine_cnt[kH];
#pragma unroll
                    for (int h = 0; h < kH; ++h) mine_cnt[h] = 0;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
#pragma unroll
                        for (int h = 0; h < kH; ++h) {
                            const uint32_t c = __popc(__ballot_sync(0xFFFFFFFFu, (uint32_t)(m >> (h * 32 + k)) & 1u));
                            if (lane == k) mine_cnt[h] = c;
                        }
                    }
#pragma unroll
                    for (int h = 0; h < kH; ++h)
                        if (mine_cnt[h]) {
                            atomicAdd(&s_msum[h * 32 + lane], (unsigned long long)mine_cnt[h] * level);
                            atomicAdd(&s_mreach[h * 32 + lane], mine_cnt[h]);
                            atomicMax(&s_mecc[h * 32 + lane], level);
                        }
                } else {
                    const uint32_t n_ent = min(32u, qend - tile);
                    for (uint32_t e = 0; e < n_ent; ++e) {
                        const uint32_t u_e = __shfl_sync(0xFFFFFFFFu, u, e);
                        const Mask m_e = __shfl_sync(0xFFFFFFFFu, m, e);
#pragma unroll
                        for (int h = 0; h < kH; ++h)
                            if ((uint32_t)(m_e >> (h * 32 + lane)) & 1u) dist[(size_t)u_e * kSrc + h * 32 + lane] = cur_d;
                    }
                }
                if (pull) continue;                  // (uniform for the whole slot) the scan below does the expansion
                const uint32_t incl = warp_inclusive_scan(deg, lane);
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const uint32_t off = incl - deg;
                // kFwdUnroll flattened entries per lane and step: every stage below is a batch of
                // independent memory operations (col -> seen/next snapshot -> atomicOr(next)).
                for (uint32_t t0 = 0; t0 < total; t0 += 32 * kFwdUnroll) {
                    uint32_t x[kFwdUnroll];
                    Mask fresh[kFwdUnroll];
#pragma unroll
                    for (int j = 0; j < kFwdUnroll; ++j) {
                        const uint32_t t = t0 + j * 32 + lane;
                        const int o = tile_owner(incl, t);
                        const uint32_t rs_o = __shfl_sync(0xFFFFFFFFu, rs, o);
                        const uint32_t off_o = __shfl_sync(0xFFFFFFFFu, off, o);
                        const Mask m_o = __shfl_sync(0xFFFFFFFFu, m, o);
                        x[j] = 0; fresh[j] = 0;
                        if (t < total) { x[j] = __ldg(p.col_idx + rs_o + (t - off_o)); fresh[j] = m_o; }
                    }
                    // every stage is a batch of kFwdUnroll INDEPENDENT operations per lane: all loads /
                    // atomics of a stage are issued before any result is consumed (predicated PTX, no
                    // branches, so the compiler cannot serialise the round trips)
                    {   // sources that already reached x, or are already noted for the level being built (a
                        // snapshot: the atomic below is exact, this only spares most of the redundant ones)
                        Mask sv[kFwdUnroll], nv[kFwdUnroll];
#pragma unroll
                        for (int j = 0; j < kFwdUnroll; ++j) {
                            sv[j] = fresh[j] ? __ldcg(&rec[(size_t)x[j] * 4]) : ~(Mask)0;
                            nv[j] = fresh[j] ? __ldcg(&next_nxt[(size_t)x[j] * 4]) : (Mask)0;
                        }
#pragma unroll
                        for (int j = 0; j < kFwdUnroll; ++j) fresh[j] &= ~(sv[j] | nv[j]);
                    }
                    uint32_t enq_bits = 0;
                    {
                        Mask old[kFwdUnroll];
#pragma unroll
                        for (int j = 0; j < kFwdUnroll; ++j) old[j] = atom_or_if(fresh[j] != 0, &next_nxt[(size_t)x[j] * 4], fresh[j]);
#pragma unroll
                        for (int j = 0; j < kFwdUnroll; ++j) enq_bits |= (fresh[j] != 0 && old[j] == 0 ? 1u : 0u) << j;
                    }
#pragma unroll
                    for (int j = 0; j < kFwdUnroll; ++j) {
                        const bool enqueue = (enq_bits >> j) & 1u;
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, enqueue);
                        if (bal) {
                            uint32_t base = 0;
                            const int leader = __ffs(bal) - 1;
                            if (lane == leader) base = tail_add((uint32_t)__popc(bal));
                            base = __shfl_sync(0xFFFFFFFFu, base, leader);
                            if (enqueue) {
                                const uint32_t pos = base + __popc(bal & lane_lt);
                                if (pos < qcap) queue_v[pos] = x[j];
                                else fail_set();
                            }
                        }
                    }
                }
            }
            if (pull) {
                // ---- pull scan: thread per vertex, vertices in id order (streams), neighbours' masks gathered ----
                uint32_t my_open = 0;
                for (uint32_t base = warp_t * 32; base < n; base += warps_t * 32) {
                    const uint32_t x = base + lane;
                    Mask unseen = 0;
                    uint32_t rs = 0, deg = 0;
                    if (x < n) {
                        unseen = ~__ldcg(&rec[(size_t)x * 4]) & valid_src;
                        if (unseen) {
                            rs = __ldg(p.row_ptr + x);
                            deg = __ldg(p.row_ptr + x + 1) - rs;
                        }
                    }
                    Mask acc = 0;
                    const bool hub = deg > kPullHubDeg;
                    if (!hub) {
                        for (uint32_t e = 0; e < deg; e += kPullUnroll) {
                            uint32_t un[kPullUnroll];
                            Mask mm[kPullUnroll];
#pragma unroll
                            for (int j = 0; j < kPullUnroll; ++j) un[j] = e + j < deg ? __ldg(p.col_idx + rs + e + j) : 0xFFFFFFFFu;
#pragma unroll
                            for (int j = 0; j < kPullUnroll; ++j) mm[j] = un[j] != 0xFFFFFFFFu ? __ldcg(&next_cur[(size_t)un[j] * 4]) : (Mask)0;
#pragma unroll
                            for (int j = 0; j < kPullUnroll; ++j) acc |= mm[j];
                            if ((acc & unseen) == unseen) break;      // nothing left to find for this vertex
                        }
                    }
                    // hubs: the whole warp gathers one adjacency list (BA: degrees up to ~ m * sqrt(N))
                    for (uint32_t hub_bal = __ballot_sync(0xFFFFFFFFu, hub); hub_bal; hub_bal &= hub_bal - 1) {
                        const int o = __ffs(hub_bal) - 1;
                        const uint32_t rs_o = __shfl_sync(0xFFFFFFFFu, rs, o), deg_o = __shfl_sync(0xFFFFFFFFu, deg, o);
                        Mask a = 0;
                        for (uint32_t e = lane; e < deg_o; e += 32) a |= __ldcg(&next_cur[(size_t)__ldg(p.col_idx + rs_o + e) * 4]);
#pragma unroll
                        for (int sh = 16; sh > 0; sh >>= 1) a |= __shfl_xor_sync(0xFFFFFFFFu, a, sh);
                        if (lane == o) acc = a;
                    }
                    const Mask fresh = acc & unseen;
                    if (fresh) __stcg(&next_nxt[(size_t)x * 4], fresh);       // one writer per vertex
                    my_open += (unseen & ~fresh) != 0 ? 1u : 0u;
                    const bool enqueue = fresh != 0;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, enqueue);
                    if (bal) {
                        uint32_t qb = 0;
                        const int leader = __ffs(bal) - 1;
                        if (lane == leader) qb = tail_add((uint32_t)__popc(bal));
                        qb = __shfl_sync(0xFFFFFFFFu, qb, leader);
                        if (enqueue) {
                            const uint32_t pos = qb + __popc(bal & lane_lt);
                            if (pos < qcap) queue_v[pos] = x;
                            else fail_set();
                        }
                    }
                }
                my_open = __reduce_add_sync(0xFFFFFFFFu, my_open);
                if (lane == 0 && my_open) open_add(my_open);
            }
            prev_pull = pull;
            if (tid_t == 0) dir_levels += pull ? (1ull << 32) : 1ull;
            ++level;
        }
        // here: levels 0..level-1 are non-empty (if !too_deep), lvl[l] = start of level l,
        // lvl[level] = end of the last level.
        team_sync();
        const bool failed = fail_read();

        if (too_deep || failed) {
            // leave the slot clean for the next batch: drop pending mask bits (rare path)
            for (uint32_t v = tid_t; v < n; v += threads_t) { __stcg(&rec[(size_t)v * 4 + 1], (Mask)0); __stcg(&rec[(size_t)v * 4 + 2], (Mask)0); }
            if (tid_t == 0) p.batch_status[batch] = failed ? kQueueOverflow : kTooDeep;
            team_sync();
            continue;
        }
        atomicAdd(&s_cnt[0], (unsigned long long)b_reached);
        atomicAdd(&s_cnt[1], (unsigned long long)b_edges);
        atomicAdd(&s_cnt[2], (unsigned long long)b_visits);
        const uint32_t depth = level - 1;             // deepest non-empty level
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_fwd += t - t_phase; t_phase = t; }
        if (kMeasure) {
            if (depth > my_depth) my_depth = depth;
            if (threadIdx.x < kSrc) {
                const uint32_t sidx = batch * kSrc + threadIdx.x;
                if (sidx < p.n_sources) {
                    if (p.out_ecc) p.out_ecc[sidx] = s_mecc[threadIdx.x];
                    if (p.out_sumdist) p.out_sumdist[sidx] = s_msum[threadIdx.x];
                    if (p.out_reached) p.out_reached[sidx] = s_mreach[threadIdx.x];
                }
            }
            if (threadIdx.x == 0) p.batch_status[batch] = kDone;
            __syncthreads();
            continue;
        }
        if (depth > my_depth) my_depth = depth;

        // ---- backward: levels depth .. 1, pull from successors --------------------------
        for (uint32_t l = depth; l >= 1; --l) {
            const uint32_t lb = lvl[l], le = lvl[l + 1];
            const uint32_t d_succ = l + 1, d_pred = l - 1;
            const uint32_t succ_tag = 0x80u | (((l + 1) & 1u) << 6), mark_tag = 0x80u | ((l & 1u) << 6);
            for (uint32_t tile = lb + warp_t * 32; tile < le; tile += warps_t * 32) {
                const uint32_t idx = tile + lane;
                uint32_t w = 0, rs = 0, deg = 0;
                Mask m = 0;
                if (idx < le) {
                    w = queue_v[idx]; m = queue_m[idx];
                    rs = __ldg(p.row_ptr + w);
                    deg = __ldg(p.row_ptr + w + 1) - rs;
                }
                const uint32_t incl = warp_inclusive_scan(deg, lane);
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const uint32_t off = incl - deg;
                // COMPACT LANES.  While queue entry (w, mask) is being consumed, lane j works for the j-th LIVE source
                // of the mask (and, for masks with more than 32 sources, also for the (j+32)-th): a batch's entries
                // carry about a quarter of the batch's sources on an expander, so one lane per source of the batch
                // would leave three lanes in four idle on every neighbour row.  The j-th set bit is found once per
                // entry (ballot-style ranks through a 64-byte list in shared memory); every neighbour row then costs
                // one distance-byte read, two compares and a predicated q load per lane.  Sums over a vertex's
                // successors are still taken by ONE lane in adjacency order: results do not depend on the mapping.
                int cur = -1;                                       // entry whose rows are being consumed (warp-uniform)
                uint32_t cur_end = 0, c_cur = 0, npred[kH], ks[kH], succ_key[kH], pred_key[kH];
                const double *qk[kH];                               // q + ks[h]: a neighbour's value sits at qk[h] + x * kSrc
                double acc[kH], my_total = 0.0;
#pragma unroll
                for (int h = 0; h < kH; ++h) { acc[h] = 0.0; npred[h] = 0; ks[h] = 0; succ_key[h] = 0xFFFFFF00u; pred_key[h] = 0xFFFFFF00u; qk[h] = q; }

                auto begin_entry = [&]() {
                    const Mask m_c = __shfl_sync(0xFFFFFFFFu, m, cur);
                    const uint32_t lo = (uint32_t)m_c;
                    uint32_t hi = 0;
                    if constexpr (kH == 2) hi = (uint32_t)(m_c >> 32);
                    const uint32_t clo = __popc(lo);
                    c_cur = clo + __popc(hi);
                    __syncwarp();                                   // the previous entry's list has been read by every lane
                    if ((lo >> lane) & 1u) s_klist[warp][__popc(lo & lane_lt)] = (unsigned char)lane;
                    if constexpr (kH == 2) {
                        if (hi != 0 && ((hi >> lane) & 1u)) s_klist[warp][clo + __popc(hi & lane_lt)] = (unsigned char)(lane + 32);
                    }
                    __syncwarp();
                    // lanes without a source compare against keys no byte (or 32-bit distance) can equal
                    const bool on0 = (uint32_t)lane < c_cur;
                    ks[0] = on0 ? (uint32_t)s_klist[warp][lane] : 0u;
                    succ_key[0] = on0 ? (kLeafMarks ? succ_tag : d_succ) : 0xFFFFFF00u;
                    pred_key[0] = on0 ? d_pred : 0xFFFFFF00u;
                    qk[0] = q + ks[0];
                    acc[0] = 0.0; npred[0] = 0;
                    if constexpr (kH == 2) {
                        const bool on1 = (uint32_t)lane + 32u < c_cur;
                        ks[1] = on1 ? (uint32_t)s_klist[warp][lane + 32] : 0u;
                        succ_key[1] = on1 ? (kLeafMarks ? succ_tag : d_succ) : 0xFFFFFF00u;
                        pred_key[1] = on1 ? d_pred : 0xFFFFFF00u;
                        qk[1] = q + ks[1];
                        acc[1] = 0.0; npred[1] = 0;
                    }
                };
                auto finish_entry = [&]() {
                    const uint32_t w_c = __shfl_sync(0xFFFFFFFFu, w, cur);
                    double contrib = 0.0;
#pragma unroll
                    for (int h = 0; h < kH; ++h) {
                        if ((uint32_t)lane + 32u * h < c_cur) {
                            const double bk = 1.0 + acc[h];
                            const size_t at = (size_t)w_c * kSrc + ks[h];
                            if constexpr (kLeafMarks) {
                                // q values are positive, so acc == 0 <=> no successor <=> bk == 1 exactly
                                const bool leaf = acc[h] == 0.0 && npred[h] <= kMaxLeafPred;
                                if (!leaf) q[at] = bk / (double)npred[h];
                                dist[at] = (DistT)(mark_tag | (leaf ? npred[h] : 0u));
                            } else q[at] = bk / (double)npred[h];
                            contrib += p.include_endpoints ? bk : bk - 1.0;
                            if (kDebug) {
                                p.dbg_npred[at] = npred[h];
                                p.dbg_bk[at] = bk;
                            }
                        }
                    }
                    const double tot = warp_sum(contrib);
                    if (lane == cur) my_total = tot;
                };
                // kRows neighbour rows of the current entry, starting at row r of the staged chunk, for kSlots source
                // slots per lane.  Straight-line code: every q load of the group is issued (predicated) before the first
                // value is consumed; rows past the group's end (nrows < kRows) re-read row 31 with their predicates off
                // and add an exact 0.0, which is cheaper than a branch per row.
                auto consume = [&](auto rows_tag, auto slots_tag, int buf, uint32_t r, uint32_t nrows) {
                    constexpr int kRows = decltype(rows_tag)::value, kSlots = decltype(slots_tag)::value;
                    double qv[kRows][kSlots];
                    const Stage *__restrict__ st = &s_stage[warp][buf][0];
                    const unsigned char *__restrict__ rows0 = &s_rows[warp][buf][0][0];
#pragma unroll
                    for (int jj = 0; jj < kRows; ++jj) {
                        const uint32_t rr = min(r + (uint32_t)jj, 31u);
                        const bool valid = (uint32_t)jj < nrows;    // warp-uniform
                        const uint32_t xj = st[rr].x;
#pragma unroll
                        for (int h = 0; h < kSlots; ++h) {
                            uint32_t b;
                            if constexpr (kStageRows) b = (uint32_t)rows0[rr * kSrc + ks[h]];
                            else b = valid ? (uint32_t)dist[(size_t)xj * kSrc + ks[h]] : 0xFFFFFF01u;
                            const double *pq = qk[h] + (size_t)xj * kSrc;
                            if constexpr (kLeafMarks) {
                                const bool succ = valid && (b & 0xC0u) == succ_key[h];
                                const uint32_t info = b & 0x3Fu;
                                qv[jj][h] = succ ? (info ? s_rcp[info] : __ldcg(pq)) : 0.0;
                            } else {
                                const bool succ = valid && b == succ_key[h];
                                qv[jj][h] = succ ? __ldcg(pq) : 0.0;
                            }
                            if (valid && b == pred_key[h]) npred[h] += 1u;
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < kRows; ++jj)              // adjacency (CSR) order: fixes the rounding
#pragma unroll
                        for (int h = 0; h < kSlots; ++h) acc[h] += qv[jj][h];
                };

                // Flattened adjacency entries are consumed in chunks of 32.  Two-deep pipeline per warp:
                // chunk s+2's neighbour ids are being loaded (registers), chunk s+1's distance rows are
                // in flight to shared memory (cp.async, 32 rows x kSrc B per chunk, 16 B pieces),
                // chunk s is consumed: one LDS byte per (row, source), then the predicated q loads.
                auto edge_of = [&](uint32_t t0) -> Stage {        // (neighbour id, owner's source mask) of flattened entry t0+lane
                    const uint32_t t = t0 + lane;
                    const int o = tile_owner(incl, t);
                    const uint32_t rs_o = __shfl_sync(0xFFFFFFFFu, rs, o);
                    const uint32_t off_o = __shfl_sync(0xFFFFFFFFu, off, o);
                    const Mask m_o = __shfl_sync(0xFFFFFFFFu, m, o);
                    return t < total ? ST::stage(__ldg(p.col_idx + rs_o + (t - off_o)), m_o) : ST::stage(0u, (Mask)0);
                };
                auto stage_rows = [&](int buf) {                  // rows of the chunk whose edges sit in s_stage[warp][buf]
                    if constexpr (kStageRows) {
                        constexpr int kPieces = kSrc / 16;        // 16 B pieces per row
#pragma unroll
                        for (int it = 0; it < kPieces; ++it) {
                            const int r = it * (32 / kPieces) + lane / kPieces;
                            const int part = lane % kPieces;
                            const Stage e = s_stage[warp][buf][r];
                            if (ST::any(e)) {
                                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_rows[warp][buf][r][part * 16]);
                                const unsigned char *src = reinterpret_cast<const unsigned char *>(dist) + (size_t)e.x * kSrc + part * 16;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                const uint32_t n_chunks = (total + 31) / 32;
                Stage st_next = edge_of(0);
                __syncwarp();
                s_stage[warp][0][lane] = st_next;
                __syncwarp();
                stage_rows(0);
                st_next = n_chunks > 1 ? edge_of(32) : ST::stage(0u, (Mask)0);
                for (uint32_t sc = 0; sc < n_chunks; ++sc) {
                    const int buf = sc & 1;
                    const uint32_t t0 = sc * 32;
                    const uint32_t cnt = min(32u, total - t0);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");   // chunk sc's rows have landed (this lane's copies)
                    __syncwarp();                                          // ... and every other lane's
                    if constexpr (kStageRows && NEC_Q_PREFETCH != 0) {
                        // The consumption below pulls q in groups of a few rows, each group one dependent DRAM round
                        // trip: the pass is bound by that latency, not by bytes or instructions.  So every q value the
                        // chunk will pull is requested up front (prefetch: no register, no scoreboard) - all 32 rows'
                        // requests are in flight together and the groups that follow find their values in L2.
                        // One lane per source of the batch here (the staged mask says which are live): no per-entry state.
                        for (uint32_t r = 0; r < cnt; ++r) {
                            const Stage e = s_stage[warp][buf][r];
#pragma unroll
                            for (int h = 0; h < kH; ++h) {
                                const uint32_t b = (uint32_t)s_rows[warp][buf][r][h * 32 + lane];
                                const bool act = (ST::half(e, h) >> lane) & 1u;
                                const bool want = act && (kLeafMarks ? (b & 0xFFu) == succ_tag : b == d_succ);   // marked leaves carry their q
                                if (want) {
                                    const double *addr = &q[(size_t)e.x * kSrc + h * 32 + lane];
                                    if (NEC_Q_PREFETCH == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
                                    else asm volatile("prefetch.global.L1 [%0];" ::"l"(addr));
                                }
                            }
                        }
                    }
                    s_stage[warp][buf ^ 1][lane] = st_next;       // chunk sc+1 (zeros past the end); buffer was drained at sc-1
                    __syncwarp();
                    stage_rows(buf ^ 1);
                    st_next = sc + 2 < n_chunks ? edge_of(t0 + 64) : ST::stage(0u, (Mask)0);
                    for (uint32_t r = 0; r < cnt;) {
                        const uint32_t tt = t0 + r;
                        if (tt >= cur_end) {                      // warp-uniform: the next entry's rows start
                            if (cur >= 0) finish_entry();
                            do { ++cur; cur_end = __shfl_sync(0xFFFFFFFFu, incl, cur); } while (tt >= cur_end);
                            begin_entry();
                        }
                        const uint32_t left = min(cur_end - tt, cnt - r);      // rows of this entry inside this chunk
                        if (kH == 2 && c_cur > 32) {
                            const uint32_t nrows = min(left, (uint32_t)(kUnroll / 2));
                            consume(std::integral_constant<int, kUnroll / 2>{}, std::integral_constant<int, kH>{}, buf, r, nrows);
                            r += nrows;
                        } else if (left > (uint32_t)(kUnroll / 2)) {
                            const uint32_t nrows = min(left, (uint32_t)kUnroll);
                            consume(std::integral_constant<int, kUnroll>{}, std::integral_constant<int, 1>{}, buf, r, nrows);
                            r += nrows;
                        } else {                                  // a short tail: the half-size group wastes fewer predicated-off rows
                            consume(std::integral_constant<int, kUnroll / 2>{}, std::integral_constant<int, 1>{}, buf, r, left);
                            r += left;
                        }
                    }
                    __syncwarp();                                 // buffer `buf` may be refilled next iteration
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                if (cur >= 0) finish_entry();
                if (idx < le && deg != 0) bslot[w] += my_total;     // one writer per vertex per level
            }
            team_sync();
        }
        if (kDebug) {
            // the sources themselves (level 0) get bk but no load contribution
            const uint32_t lb = lvl[0], le = lvl[1];
            for (uint32_t idx = lb + warp; idx < le; idx += kWarps) {
                const uint32_t ex = queue_v[idx];
                const bool active = (uint32_t)(queue_m[idx] >> lane) & 1u;
                const uint32_t rs = p.row_ptr[ex], re = p.row_ptr[ex + 1];
                double acc = 0.0;
                for (uint32_t e = rs; e < re; ++e) {
                    const uint32_t x = p.col_idx[e];
                    if (active && (uint32_t)dist[(size_t)x * kSrc + lane] == 1u) acc += q[(size_t)x * kSrc + lane];
                }
                if (active) {
                    p.dbg_npred[(size_t)ex * kSrc + lane] = 0;
                    p.dbg_bk[(size_t)ex * kSrc + lane] = 1.0 + acc;
                    if (p.dbg_sigma) p.dbg_sigma[(size_t)ex * kSrc + lane] = 1.0;
                }
            }
            // SIGMA: f64 shortest-path counts, levels ascending over the same queues, pull form like the backward
            // pass (warp per queue entry, lane per source, one writer per (vertex, source), neighbours in CSR order):
            // sigma(x) = sum of sigma(u) over the neighbours one level closer.  Counts are integers, exact in f64
            // below 2^53 whatever the order; beyond that they round (nec_bfs_debug).  The load never reads them.
            if (p.dbg_sigma) {
                for (uint32_t l = 1; l <= depth; ++l) {
                    __syncthreads();
                    const uint32_t b0 = lvl[l], b1 = lvl[l + 1];
                    for (uint32_t idx = b0 + warp; idx < b1; idx += kWarps) {
                        const uint32_t ex = queue_v[idx];
                        const bool active = (uint32_t)(queue_m[idx] >> lane) & 1u;
                        const uint32_t rs = p.row_ptr[ex], re = p.row_ptr[ex + 1];
                        double acc = 0.0;
                        for (uint32_t e = rs; e < re; ++e) {
                            const uint32_t u = p.col_idx[e];
                            if (active && (uint32_t)dist[(size_t)u * kSrc + lane] == l - 1) acc += p.dbg_sigma[(size_t)u * kSrc + lane];
                        }
                        if (active) p.dbg_sigma[(size_t)ex * kSrc + lane] = acc;
                    }
                }
            }
        }
        if (threadIdx.x == 0) { if (team_rank == 0) p.batch_status[batch] = kDone; cyc_bwd += clock64() - t_phase; }
        team_sync();
    }

    // ---- flush counters -------------------------------------------------------------------
    team_sync();                                         // (rank 0's shared memory stays alive until every CTA of the team is done)
    if (threadIdx.x < 3) atomicAdd(&p.counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x == 0) {
        atomicMax(&p.counters[3], (unsigned long long)my_depth);
        atomicAdd(&p.counters[4], (unsigned long long)cyc_reset);
        atomicAdd(&p.counters[5], (unsigned long long)cyc_fwd);
        atomicAdd(&p.counters[6], (unsigned long long)cyc_bwd);
        if (dir_levels) atomicAdd(&p.counters[7], dir_levels);
    }
}

// out[v] = sum over slots (ascending) of the slot-private partial loads.
__global__ void reduce_slots_kernel(const double *__restrict__ bslot_base, size_t bslot_stride,
                                    uint32_t n_slots, uint32_t n, double *__restrict__ out) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (uint32_t s = 0; s < n_slots; ++s) acc += bslot_base[(size_t)s * bslot_stride + v];
        out[v] = acc;
    }
}

// The same for SMALL n and many slots (one source per CTA on a graph of a thousand vertices: a thousand slots, four
// blocks' worth of vertices): 32 vertices per block, warp g sums the slots g, g + G, g + 2G, ... in ascending order,
// the G partial sums are added in group order.  The grouping depends on nothing but G: reproducible run to run.
template <int G>
__global__ void __launch_bounds__(32 * G) reduce_slots_grouped_kernel(const double *__restrict__ bslot_base, size_t bslot_stride,
                                                                      uint32_t n_slots, uint32_t n, double *__restrict__ out) {
    __shared__ double part[G][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (uint32_t v0 = blockIdx.x * 32; v0 < n; v0 += gridDim.x * 32) {
        const uint32_t v = v0 + lane;
        double acc = 0.0;
        if (v < n) {
            const double *col = bslot_base + v;
            uint32_t s = grp;
            for (; s + 3 * G < n_slots; s += 4 * G) {           // four loads in flight
                const double a0 = col[(size_t)s * bslot_stride], a1 = col[(size_t)(s + G) * bslot_stride];
                const double a2 = col[(size_t)(s + 2 * G) * bslot_stride], a3 = col[(size_t)(s + 3 * G) * bslot_stride];
                acc += a0; acc += a1; acc += a2; acc += a3;
            }
            for (; s < n_slots; s += G) acc += col[(size_t)s * bslot_stride];
        }
        part[grp][lane] = acc;
        __syncthreads();
        if (grp == 0 && v < n) {
            double t = 0.0;
#pragma unroll
            for (int g = 0; g < G; ++g) t += part[g][lane];
            out[v] = t;
        }
        __syncthreads();
    }
}

// a[v] += b[v]
__global__ void add_into_kernel(double *__restrict__ a, const double *__restrict__ b, uint32_t n) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) a[v] += b[v];
}

// Mid engine's adjacency slab and degree bytes from the device CSR (nec_graph upload): row v =
// [degree, first W-1 neighbours, zero padding]; one thread per slab word, so the stores are coalesced.
__global__ void build_slab_kernel(uint32_t n, int W, const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_idx,
                                  uint32_t *__restrict__ slab, uint8_t *__restrict__ deg8) {
    const size_t words = (size_t)n * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = (uint32_t)(i / W), k = (uint32_t)(i % W);
        const uint32_t rs = row_ptr[v], deg = row_ptr[v + 1] - rs;
        uint32_t w = 0;
        if (k == 0) { w = deg; deg8[v] = (uint8_t)(deg < 255u ? deg : 255u); }
        else if (k - 1 < deg) w = col_idx[rs + k - 1];
        slab[i] = w;
    }
}

// Distance measures of the sources one engine handed to the other, scattered back to their places.
__global__ void scatter_measures_kernel(uint32_t k, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ pe,
                                        const unsigned long long *__restrict__ ps, const uint32_t *__restrict__ pr,
                                        uint32_t *__restrict__ ecc, unsigned long long *__restrict__ sum, uint32_t *__restrict__ reached) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
        const uint32_t at = idx[j];
        ecc[at] = pe[j]; sum[at] = ps[j]; reached[at] = pr[j];
    }
}

__global__ void fill_f64_kernel(double *__restrict__ a, size_t count, double value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) a[i] = value;
}

// 8-bit "unreached" (0xFF) -> 32-bit "unreached" (0xFFFFFFFF) after widening (nec_bfs_debug)
__global__ void widen_inf_kernel(uint32_t *__restrict__ d, uint32_t n) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
        if (d[v] == 0xFFu) d[v] = 0xFFFFFFFFu;
}

// Extract lane `lane` of a source-minor plane into a dense vector (nec_bfs_debug).
template <typename T, typename U>
__global__ void extract_lane_kernel(const T *__restrict__ plane, uint32_t n, int lane, U *__restrict__ out) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
        out[v] = (U)plane[(size_t)v * kLanes + lane];
}

}  // namespace nec
