// Compact batched engine ("wide" engine) for large graphs: vertex_load of B = 64, 128 or 256 sources per batch with
// the backward pass working on COMPACT per-entry state instead of source-minor rows.
//
// Same quantity as nec_kernels.cuh (reference: src/generic_graph/generic_graph.rs:1310-1385, pull form):
//   bk_s(w) = 1 + sum_{x: w in P_s(x)} bk_s(x)/|P_s(x)|,   b[v] += bk_s(v) (- 1 if !include_endpoints).
//
// Why another layout.  ncu on the source-minor engine (profiles/r2_batched_compact_summary.txt) shows the GPU
// moving 4.5 TB/s of random 32-byte sectors, each costing a 64-byte DRAM burst: per (queue entry, neighbour) row a
// 64 B distance row plus ~4 scattered sectors of the 512 B q row of the neighbour, and q stores that fill a third
// of every sector they dirty.  Here
//   * a queue entry (vertex w on level l with source mask M) owns a CONTIGUOUS run of popc(M) doubles in `qc`
//     (position assigned by a prefix sum when the level is prepared): q stores are dense and coalesced, and a
//     puller finds the values of a neighbour's sources packed in a few adjacent 64 B blocks;
//   * what a puller needs to know about neighbour x is not x's distance row but two masks: M_{l+1}(x) (sources
//     for which x is one level deeper: successors; their rank in the mask is the index into x's run) and
//     M_{l-1}(x) (predecessors, counted).  They live in a per-vertex LEVEL RECORD of four slots (level mod 4,
//     each {mask, position, level tag}), arranged so that the slots of levels l-1 and l+1 share one aligned
//     block of B / 2 bytes: one or two DRAM bursts per row, whatever the number of live sources;
//   * the forward pass writes no distance rows at all (levels are recorded by the queues only), which also
//     removes the depth limit of byte distances;
//   * wide batches: the forward pass and the row traffic are per batch, so every doubling of B halves them per
//     source, and entries of the widest level fill the lanes (~70 of 128, ~160 of 256 sources).
// Lanes are compact (lane j = j-th live source of the entry being consumed, up to four per lane = 128 sources; at
// B = 256 an entry with more live sources is consumed as two work items: ranks 0..127 and the rest), one lane sums
// one (vertex, source) value over the neighbours in adjacency order: results do not depend on the mapping, on
// positions or on the team size, and are bit-reproducible run to run.
//
// A slot (one batch in flight) is run by a TEAM of CTAs which share the queue tail, the position counter and the
// failure flag and meet at team barriers: either consecutive CTAs of a cooperative launch with those scalars and a
// barrier counter in global memory (kSoft, the default: any team size, every SM usable), or the CTAs of one
// thread-block cluster with the scalars in rank 0's shared memory (clusters are placed per GPC: only 45 clusters of
// three 1024-thread CTAs fit on 148 SMs).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "nec_kernels.cuh"

// Rows (entry, neighbour) a warp stages and consumes per step, and the bytes added to a staged row's pitch in shared memory.
// Lane r cooks row r: at a pitch of 128 B all 32 lanes hit the same banks; at 144 B a quarter warp is conflict-free
// (112.7 -> 126.8 GTEPS with 28 rows per chunk in the room of 32 unpadded ones).  The chunk size also sets the L1 size, and both
// passes live on L1 (it bounds the misses an SM keeps in flight): a CTA under 164 KB of shared memory leaves 92 KB of L1 instead
// of 60, which makes the forward pass 18 % and the backward pass 7 % faster at equal chunk size, while every row taken from
// the chunk costs the backward pass ~0.4 % (more chunks per tile).  25 rows is the most that fits under 164 KB after the
// kernel's static shared memory was cut from 33 to 17 KB: 1270 -> 1249 ms per 37 888 sources against 21 rows with the old
// layout, 1291 ms at 26 rows (over the step: 60 KB of L1).  32 padded rows (28 KB of L1) slow the forward pass down 1.6 x.
// (What in the forward pass needs the L1 is not the pull scan's adjacency rows: copying a warp's contiguous span of col_idx
// into shared memory with coalesced loads left the dependence as it was and cost 1.3 %.  Not spills either: the 139 LDL/STL
// of this kernel save batch-level values around the phases, none sits in an inner loop - scripts/sass_lines.py.)
// (At B = 128 the two staging buffers leave no such trade.)
#ifndef NEC_WIDE_CHUNK256
#define NEC_WIDE_CHUNK256 25
#endif
#ifndef NEC_WIDE_PAD256
#define NEC_WIDE_PAD256 16
#endif
// A/B knobs (scripts/exp_r2.py --lib), both measured and left off (profiles/r2_ab_wide_engine.log):
//  * two row buffers per warp at B = 256 (chunk sc+1 in flight while chunk sc is consumed) with chunks of half the size:
//    10 % slower - the time goes into the q pulls of a chunk's row groups, not into waiting for its rows, and every chunk
//    has a fixed cost;
//  * staging a chunk with ONE bulk copy per row (cp.async.bulk, 128 B, lane r requests row r, completion counted by a
//    per-warp mbarrier) instead of eight 16-byte cp.async pieces per row spread over the lanes: 2 % slower.
#ifndef NEC_WIDE_BUFS256
#define NEC_WIDE_BUFS256 1
#endif
#ifndef NEC_WIDE_BULK
#define NEC_WIDE_BULK 0
#endif

namespace nec {

struct WideParams {
    uint32_t n;
    const uint32_t *__restrict__ row_ptr;
    const uint32_t *__restrict__ col_idx;
    const uint32_t *__restrict__ sources;
    uint32_t n_sources;
    const uint32_t *__restrict__ work_list;   // batch ids to run (nullptr: 0..n_work-1)
    uint32_t n_work;
    int include_endpoints;
    uint32_t pull_mode;                        // 0 push only, 1 cost model, 2 pull only
    uint32_t avg_deg;
    // per-slot arenas (slot = cluster index)
    uint8_t *rec_base;   size_t rec_stride;    // bytes: n forward records {seen, next[even], next[odd], claim}
    uint8_t *lr_base;    size_t lr_stride;     // bytes: n level records (4 slots {mask, position, tag})
    double *qc_base;     size_t qc_stride;     // doubles: compact q values, n * B per slot
    double *bslot_base;  size_t bslot_stride;  // doubles per slot (n)
    uint32_t *queue_v_base;                    // level queues: vertex ids,
    uint8_t *queue_m_base;                     // their source masks (B / 8 bytes per entry)
    uint32_t *queue_p_base;                    // and the position of the entry's run in qc
    size_t queue_stride;                       // entries per slot (pitch)
    size_t queue_cap;                          // usable entries per slot
    uint32_t *lvl_base;  size_t lvl_stride;    // u32 per slot (max_levels + 2)
    uint32_t max_levels;
    // 256 sources per batch only: the backward pass's work list (entries with more than 128 live sources appear twice:
    // ranks 0..127 and 128..) with its per-level item counts, and where the second halves' partial load starts
    uint32_t *bq_base;   size_t bq_stride;     // u32 per slot (2 x queue pitch)
    uint32_t *bcnt_base;                       // u32 per slot (lvl_stride)
    size_t bslot_hi_off;                       // doubles
    uint8_t *batch_status;                     // per batch id (BatchStatus)
    unsigned long long *counters;              // as EngineParams::counters
    // software teams (kSoft): CTAs blockIdx.x in [slot * team_size, (slot + 1) * team_size) run slot `slot`; their shared
    // scalars and barrier counter live in global memory (8 u32 per slot: tail, fail, open, pos, barrier, -, -, -)
    uint32_t team_size;
    uint32_t *team_base;
    // != 0: stop after the forward pass and only count the queue entries (counters[2]): the planner's probe of how many distinct
    // levels a vertex sits on within a batch (nec_api.cu, engine_hint); nothing is added to the partial loads
    int forward_only;
};

struct alignas(16) WideVec256 { uint4 a, b; };
template <int kW> struct WideTraits;
template <> struct WideTraits<128> {
    using Vec = uint4;
    static __device__ __forceinline__ void unpack(const Vec &v, uint32_t (&w)[4]) { w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    static __device__ __forceinline__ Vec pack(const uint32_t (&w)[4]) { return make_uint4(w[0], w[1], w[2], w[3]); }
    static __device__ __forceinline__ Vec ldcg(const Vec *p) { return __ldcg(p); }
    static __device__ __forceinline__ void stcg(Vec *p, const Vec &v) { __stcg(p, v); }
};
template <> struct WideTraits<64> {
    using Vec = uint2;
    static __device__ __forceinline__ void unpack(const Vec &v, uint32_t (&w)[2]) { w[0] = v.x; w[1] = v.y; }
    static __device__ __forceinline__ Vec pack(const uint32_t (&w)[2]) { return make_uint2(w[0], w[1]); }
    static __device__ __forceinline__ Vec ldcg(const Vec *p) { return __ldcg(p); }
    static __device__ __forceinline__ void stcg(Vec *p, const Vec &v) { __stcg(p, v); }
};
template <> struct WideTraits<256> {
    using Vec = WideVec256;
    static __device__ __forceinline__ void unpack(const Vec &v, uint32_t (&w)[8]) {
        w[0] = v.a.x; w[1] = v.a.y; w[2] = v.a.z; w[3] = v.a.w; w[4] = v.b.x; w[5] = v.b.y; w[6] = v.b.z; w[7] = v.b.w;
    }
    static __device__ __forceinline__ Vec pack(const uint32_t (&w)[8]) {
        Vec v; v.a = make_uint4(w[0], w[1], w[2], w[3]); v.b = make_uint4(w[4], w[5], w[6], w[7]); return v;
    }
    static __device__ __forceinline__ Vec ldcg(const Vec *p) { Vec v; v.a = __ldcg(&p->a); v.b = __ldcg(&p->b); return v; }
    static __device__ __forceinline__ void stcg(Vec *p, const Vec &v) { __stcg(&p->a, v.a); __stcg(&p->b, v.b); }
};

// Shared-memory layout (dynamic): per warp kBufs buffers of 32 staged rows (kW / 2 bytes each; two buffers, one at 256 sources
// per batch where two would not fit), one group of rows of read-ahead padding, then the masks of each warp's tile of entries.
template <int kW> __host__ __device__ constexpr int wide_row_bufs() { return kW == 256 ? NEC_WIDE_BUFS256 : 2; }
template <int kW> __host__ __device__ constexpr int wide_chunk_rows() { return kW == 256 ? NEC_WIDE_CHUNK256 : 32; }
template <int kW> __host__ __device__ constexpr int wide_row_pitch() { return kW / 2 + (kW == 256 ? NEC_WIDE_PAD256 : 0); }
template <int kW, int kThreads>
constexpr size_t wide_smem_bytes() {
    return (size_t)(kThreads / 32) * wide_row_bufs<kW>() * wide_chunk_rows<kW>() * wide_row_pitch<kW>() + 8 * wide_row_pitch<kW>() + (size_t)(kThreads / 32) * 32 * (kW / 8);
}

constexpr uint32_t kWideInvalid = 0xFFFFFFFFu;

// kSoft = false: a slot's CTAs are one thread-block cluster (shared scalars in rank 0's shared memory, cluster barriers).
// kSoft = true : a slot's CTAs are any `team_size` consecutive CTAs of a cooperative launch (all CTAs co-resident); the
//                shared scalars and a barrier counter live in global memory.  Clusters are placed per GPC, so only 45
//                clusters of three 1024-thread CTAs (or 33 of four) fit on the 148 SMs; software teams use all of them.
template <int kW, int kThreads, bool kSoft = false>
__global__ void __launch_bounds__(kThreads, kThreads == 512 ? 2 : 1)
wide_engine(const WideParams p) {
    namespace cg = cooperative_groups;
    using WT = WideTraits<kW>;
    using Vec = typename WT::Vec;
    constexpr int kWords = kW / 32;                     // mask words
    constexpr int kLaneSlots = kWords < 4 ? kWords : 4; // the most sources a lane works for at once (128 per pass of an entry)
    constexpr bool kVirtual = kW == 256;                // entries with more than 128 live sources are consumed in two passes
    constexpr int kBufs = wide_row_bufs<kW>();
    constexpr int kPullU = kW == 256 ? 2 : kPullUnroll; // neighbours' frontier masks in flight per thread (forward, pull)
    constexpr int kVecB = kW / 8;                       // bytes of a mask
    constexpr int kRecB = kW / 2;                       // forward record: seen, next[2], claim (+ padding)
    constexpr int kSlotB = kW / 4;                      // level-record slot: mask, position, tag (+ padding at B = 128)
    constexpr int kRowB = kW / 2;                       // staged row: the two slots of levels l-1 / l+1 = kWords cooked 16 B cells
    constexpr int kPieces = kRowB / 16;                 // 16 B pieces per row
    constexpr uint32_t kChunk = wide_chunk_rows<kW>();  // rows per chunk (<= 32: lane r stages and cooks row r)
    constexpr int kRowS = wide_row_pitch<kW>();     // row pitch in shared memory: lane r cooks row r, and at a pitch of 64 / 128 B all
                                                        // 32 lanes would hit the same banks (16- / 32-way conflicts on every access)
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t team_size, team_rank;
    if constexpr (kSoft) { team_size = p.team_size; team_rank = blockIdx.x % team_size; }
    else { team_size = cg::this_cluster().num_blocks(); team_rank = cg::this_cluster().block_rank(); }
    const uint32_t slot = blockIdx.x / team_size, n_slots = gridDim.x / team_size;
    const uint32_t tid_t = team_rank * kThreads + threadIdx.x, threads_t = team_size * kThreads;
    const uint32_t warp_t = team_rank * kWarps + warp, warps_t = team_size * kWarps;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t n = p.n;

    uint8_t *__restrict__ rec = p.rec_base + (size_t)slot * p.rec_stride;
    uint8_t *__restrict__ lr = p.lr_base + (size_t)slot * p.lr_stride;
    double *__restrict__ qc = p.qc_base + (size_t)slot * p.qc_stride;
    double *__restrict__ bslot = p.bslot_base + (size_t)slot * p.bslot_stride;
    uint32_t *__restrict__ queue_v = p.queue_v_base + (size_t)slot * p.queue_stride;
    Vec *__restrict__ queue_m = reinterpret_cast<Vec *>(p.queue_m_base + (size_t)slot * p.queue_stride * kVecB);
    uint32_t *__restrict__ queue_p = p.queue_p_base + (size_t)slot * p.queue_stride;
    uint32_t *__restrict__ lvl = p.lvl_base + (size_t)slot * p.lvl_stride;
    const uint32_t qcap = (uint32_t)p.queue_cap;
    // the slot's q base lives in a register pair: without this the compiler re-derives it from the kernel
    // parameters (two constant loads, a 64-bit multiply-add, a 64-bit shift-add) in front of every one of the
    // hot loop's loads
    const double *qc_hot = qc;
    asm volatile("" : "+l"(qc_hot));

    auto seen_of = [&](uint32_t v) -> Vec * { return reinterpret_cast<Vec *>(rec + (size_t)v * kRecB); };
    auto next_of = [&](uint32_t v, uint32_t parity) -> Vec * { return reinterpret_cast<Vec *>(rec + (size_t)v * kRecB + kVecB * (1 + parity)); };
    auto claim_of = [&](uint32_t v) -> uint32_t * { return reinterpret_cast<uint32_t *>(rec + (size_t)v * kRecB + 3 * kVecB); };

    // team-shared scalars: rank 0's copies, reached through distributed shared memory
    __shared__ uint32_t s_tail, s_fail, s_open, s_pos;
    uint32_t *tail_p, *fail_p, *open_p, *pos_p, *bar_p = nullptr;
    if constexpr (kSoft) {
        uint32_t *t = p.team_base + (size_t)slot * 8;
        tail_p = t; fail_p = t + 1; open_p = t + 2; pos_p = t + 3; bar_p = t + 4;
    } else {
        tail_p = cg::this_cluster().map_shared_rank(&s_tail, 0);
        fail_p = cg::this_cluster().map_shared_rank(&s_fail, 0);
        open_p = cg::this_cluster().map_shared_rank(&s_open, 0);
        pos_p = cg::this_cluster().map_shared_rank(&s_pos, 0);
    }
    // team barrier.  Software form (as cooperative groups' grid barrier, per team): the CTA's writes are ordered before
    // thread 0's release by the block barrier, thread 0 takes a ticket on the team's counter and spins with acquire loads
    // until the whole team has arrived, the second block barrier orders everything after it.
    auto team_sync = [&]() {
        if constexpr (kSoft) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                const uint32_t ticket = atomicAdd(bar_p, 1u);
                const uint32_t target = (ticket / team_size + 1u) * team_size;
                uint32_t seen;
                do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar_p) : "memory"); } while ((int32_t)(seen - target) < 0);
                __threadfence();
            }
            __syncthreads();
        } else
            cg::this_cluster().sync();
    };
    // Static shared memory is kept small on purpose: together with the staging buffers it decides how much L1 the SM keeps.
    __shared__ unsigned char s_klist[kWarps][kW > 128 ? 128 : kW];   // backward: the live sources of the work item a warp is consuming (at most 128)
    __shared__ uint32_t s_tile_p[kWarps][32];             // backward: run positions of the warp's tile of entries (masks: dynamic, below)
    __shared__ uint32_t s_tile_base[kWarps][32], s_tile_end[kWarps][32];   // per entry: adjacency row start minus its first flattened row, and its past-the-end flattened row
    // (an entry's contribution to the partial load is written over its mask in my_tile_m once the entry has been begun)
    __shared__ unsigned long long s_cnt[3];
    constexpr bool kBulk = NEC_WIDE_BULK != 0 && kW == 256 && wide_row_bufs<kW>() == 1;
    __shared__ __align__(8) unsigned long long s_mbar[kBulk ? kWarps : 1];   // bulk staging: a warp's mbarrier ...
    __shared__ uint32_t s_mphase[kBulk ? kWarps : 1];                        // ... and the parity of its next phase
    extern __shared__ __align__(16) unsigned char wide_smem[];
    unsigned char *my_rows = wide_smem + (size_t)warp * kBufs * kChunk * kRowS;   // [kBufs][kChunk][kRowS]
    uint32_t (*my_tile_m)[kWords] = reinterpret_cast<uint32_t (*)[kWords]>(wide_smem + (size_t)kWarps * kBufs * kChunk * kRowS + 8 * kRowS) + (size_t)warp * 32;   // [32][kWords]
    uint32_t *__restrict__ bq = kVirtual ? p.bq_base + (size_t)slot * p.bq_stride : nullptr;
    uint32_t *__restrict__ bcnt = kVirtual ? p.bcnt_base + (size_t)slot * p.lvl_stride : nullptr;

    if constexpr (kBulk) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[warp])) : "memory");
            s_mphase[warp] = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    uint32_t my_depth = 0;
    long long cyc_reset = 0, cyc_fwd = 0, cyc_bwd = 0;      // thread 0 only
    unsigned long long dir_levels = 0;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;

    for (uint32_t work = slot; work < p.n_work; work += n_slots) {
        const uint32_t batch = p.work_list ? p.work_list[work] : work;
        long long t_phase = clock64();
        // ---- reset: forward records and level records (tags) of the slot ------------------------------------
        {
            uint4 *r16 = reinterpret_cast<uint4 *>(rec);
            const size_t n16 = (size_t)n * kRecB / 16;
            for (size_t i = tid_t; i < n16; i += threads_t) __stcg(&r16[i], make_uint4(0u, 0u, 0u, 0u));
            uint4 *l16 = reinterpret_cast<uint4 *>(lr);
            const size_t m16 = (size_t)n * 4 * kSlotB / 16;
            for (size_t i = tid_t; i < m16; i += threads_t) __stcg(&l16[i], make_uint4(0u, 0u, 0u, 0u));
        }
        if constexpr (kSoft) { if (tid_t == 0) { *(volatile uint32_t *)tail_p = 0; *(volatile uint32_t *)fail_p = 0; *(volatile uint32_t *)open_p = 0; *(volatile uint32_t *)pos_p = 0; } }
        else if (threadIdx.x == 0) { s_tail = 0; s_fail = 0; s_open = 0; s_pos = 0; }
        team_sync();

        // ---- seed level 0 -----------------------------------------------------------------------------------
        uint32_t valid_src[kWords];
        {
            const uint32_t first = batch * kW;
            const uint32_t have = p.n_sources - first < (uint32_t)kW ? p.n_sources - first : (uint32_t)kW;
#pragma unroll
            for (int wd = 0; wd < kWords; ++wd) {
                const uint32_t lo = wd * 32;
                valid_src[wd] = have >= lo + 32 ? 0xFFFFFFFFu : (have > lo ? (1u << (have - lo)) - 1u : 0u);
            }
            if (warp_t == 0) {
#pragma unroll
                for (int wd = 0; wd < kWords; ++wd) {
                    if ((valid_src[wd] >> lane) & 1u) {
                        const uint32_t s = p.sources[first + wd * 32 + lane];
                        atomicOr(reinterpret_cast<uint32_t *>(next_of(s, 0)) + wd, 1u << lane);
                        if (atomicExch(claim_of(s), 1u) != 1u) queue_v[atomicAdd(tail_p, 1u)] = s;
                    }
                }
            }
        }
        team_sync();
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_reset += t - t_phase; t_phase = t; }

        // ---- forward: level-synchronous, bit-parallel, direction chosen per level ---------------------------
        uint32_t level = 0, qbeg = 0, qend = 0;
        bool too_deep = false;
        uint32_t n_open = n, defer_beg = 0, defer_end = 0;
        bool prev_pull = false;
        for (;;) {
            team_sync();                              // previous level fully expanded
            const uint32_t tail_raw = *(volatile uint32_t *)tail_p;
            const uint32_t open_raw = *(volatile uint32_t *)open_p;
            const uint32_t tail_now = tail_raw < qcap ? tail_raw : qcap;
            team_sync();                              // every warp sampled the tail before it moves again
            qbeg = qend;
            qend = tail_now;
            if (prev_pull) n_open = open_raw;
            if (tid_t == 0) lvl[level] = qbeg;
            if (qend == qbeg) break;
            if (level >= p.max_levels) { too_deep = true; break; }
            const uint32_t par_c = level & 1u, par_n = par_c ^ 1u;
            const uint32_t entries = qend - qbeg;
            const bool pull = p.pull_mode == 2 ||
                              (p.pull_mode == 1 && 2ull * entries * p.avg_deg > (9ull * n) / 16 + (unsigned long long)n_open * p.avg_deg);
            // mark: the level's vertices become `seen` for their sources before anything is expanded; the entry's
            // mask moves to the queue and next[] cleans itself (one level later if this level is pulled: the scan
            // reads the masks from the records)
            for (uint32_t idx = qbeg + tid_t; idx < qend; idx += threads_t) {
                const uint32_t u = queue_v[idx];
                const Vec m = WT::ldcg(next_of(u, par_c));
                const Vec sn = WT::ldcg(seen_of(u));
                uint32_t mw[kWords], sw[kWords];
                WT::unpack(m, mw); WT::unpack(sn, sw);
#pragma unroll
                for (int wd = 0; wd < kWords; ++wd) sw[wd] |= mw[wd];
                WT::stcg(seen_of(u), WT::pack(sw));
                if (!pull) {
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) sw[wd] = 0;
                    WT::stcg(next_of(u, par_c), WT::pack(sw));
                }
                queue_m[idx] = m;
            }
            {
                uint32_t zero[kWords];
#pragma unroll
                for (int wd = 0; wd < kWords; ++wd) zero[wd] = 0;
                for (uint32_t idx = defer_beg + tid_t; idx < defer_end; idx += threads_t)
                    WT::stcg(next_of(queue_v[idx], par_n), WT::pack(zero));     // the pulled level-1's masks: same parity as level+1
            }
            defer_beg = pull ? qbeg : 0u;
            defer_end = pull ? qend : 0u;
            if constexpr (kSoft) { if (pull && tid_t == 0) *(volatile uint32_t *)open_p = 0; }
            else if (pull && threadIdx.x == 0) s_open = 0;
            team_sync();

            if (!pull) {
                // ---- push: one lane per adjacency entry of the frontier ----
                const uint32_t tag = level + 2u;         // claim tag of the level being built
                for (uint32_t tile = qbeg + warp_t * 32; tile < qend; tile += warps_t * 32) {
                    const uint32_t idx = tile + lane;
                    uint32_t m[kWords], rs = 0, deg = 0;
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) m[wd] = 0;
                    if (idx < qend) {
                        const uint32_t u = queue_v[idx];
                        WT::unpack(queue_m[idx], m);
                        rs = __ldg(p.row_ptr + u);
                        deg = __ldg(p.row_ptr + u + 1) - rs;
                    }
                    const uint32_t incl = warp_inclusive_scan(deg, lane);
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    const uint32_t off = incl - deg;
                    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
                        const uint32_t t = t0 + lane;
                        const int o = tile_owner(incl, t);
                        const uint32_t rs_o = __shfl_sync(0xFFFFFFFFu, rs, o);
                        const uint32_t off_o = __shfl_sync(0xFFFFFFFFu, off, o);
                        uint32_t fresh[kWords];
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) fresh[wd] = __shfl_sync(0xFFFFFFFFu, m[wd], o);
                        bool enqueue = false;
                        uint32_t x = 0;
                        if (t < total) {
                            x = __ldg(p.col_idx + rs_o + (t - off_o));
                            uint32_t sv[kWords], nv[kWords];
                            WT::unpack(WT::ldcg(seen_of(x)), sv);
                            WT::unpack(WT::ldcg(next_of(x, par_n)), nv);
                            bool cand = false;
#pragma unroll
                            for (int wd = 0; wd < kWords; ++wd) {
                                fresh[wd] &= ~(sv[wd] | nv[wd]);
                                if (fresh[wd]) cand |= atomicOr(reinterpret_cast<uint32_t *>(next_of(x, par_n)) + wd, fresh[wd]) == 0u;
                            }
                            // several lanes can each be first on a different word: the claim decides who enqueues x
                            if (cand) enqueue = atomicExch(claim_of(x), tag) != tag;
                        }
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, enqueue);
                        if (bal) {
                            uint32_t base = 0;
                            const int leader = __ffs(bal) - 1;
                            if (lane == leader) base = atomicAdd(tail_p, (uint32_t)__popc(bal));
                            base = __shfl_sync(0xFFFFFFFFu, base, leader);
                            if (enqueue) {
                                const uint32_t pos = base + __popc(bal & lane_lt);
                                if (pos < qcap) queue_v[pos] = x;
                                else *(volatile uint32_t *)fail_p = 1;
                            }
                        }
                    }
                }
            } else {
                // ---- pull scan: thread per vertex (id order: streams), neighbours' frontier masks gathered ----
                uint32_t my_open = 0;
                for (uint32_t base = warp_t * 32; base < n; base += warps_t * 32) {
                    const uint32_t x = base + lane;
                    uint32_t unseen[kWords], acc[kWords];
                    uint32_t any_unseen = 0;
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) { unseen[wd] = 0; acc[wd] = 0; }
                    uint32_t rs = 0, deg = 0;
                    if (x < n) {
                        WT::unpack(WT::ldcg(seen_of(x)), unseen);
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) { unseen[wd] = ~unseen[wd] & valid_src[wd]; any_unseen |= unseen[wd]; }
                        if (any_unseen) {
                            rs = __ldg(p.row_ptr + x);
                            deg = __ldg(p.row_ptr + x + 1) - rs;
                        }
                    }
                    const bool hub = deg > kPullHubDeg;
                    if (!hub) {
                        // the vertex's neighbour ids are read eight at a time - one pass over the L1 lines its row shares with
                        // the rows of the neighbouring threads - and their frontier masks kPullU at a time
                        bool more = true;
                        for (uint32_t e = 0; e < deg && more; e += 8) {
                            uint32_t un[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) un[j] = e + j < deg ? __ldg(p.col_idx + rs + e + j) : kWideInvalid;
#pragma unroll
                            for (int j0 = 0; j0 < 8; j0 += kPullU) {
                                if (!more) break;
                                Vec mm[kPullU];
#pragma unroll
                                for (int j = 0; j < kPullU; ++j) {
                                    uint32_t z[kWords];
#pragma unroll
                                    for (int wd = 0; wd < kWords; ++wd) z[wd] = 0;
                                    mm[j] = un[j0 + j] != kWideInvalid ? WT::ldcg(next_of(un[j0 + j], par_c)) : WT::pack(z);
                                }
                                uint32_t missing = 0;
#pragma unroll
                                for (int j = 0; j < kPullU; ++j) {
                                    uint32_t w4[kWords];
                                    WT::unpack(mm[j], w4);
#pragma unroll
                                    for (int wd = 0; wd < kWords; ++wd) acc[wd] |= w4[wd];
                                }
#pragma unroll
                                for (int wd = 0; wd < kWords; ++wd) missing |= unseen[wd] & ~acc[wd];
                                more = missing != 0 && e + j0 + kPullU < deg;      // nothing left to find for this vertex, or no neighbour left
                            }
                        }
                    }
                    // hubs: the whole warp gathers one adjacency list
                    for (uint32_t hub_bal = __ballot_sync(0xFFFFFFFFu, hub); hub_bal; hub_bal &= hub_bal - 1) {
                        const int o = __ffs(hub_bal) - 1;
                        const uint32_t rs_o = __shfl_sync(0xFFFFFFFFu, rs, o), deg_o = __shfl_sync(0xFFFFFFFFu, deg, o);
                        uint32_t a[kWords];
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) a[wd] = 0;
                        for (uint32_t e = lane; e < deg_o; e += 32) {
                            uint32_t w4[kWords];
                            WT::unpack(WT::ldcg(next_of(__ldg(p.col_idx + rs_o + e), par_c)), w4);
#pragma unroll
                            for (int wd = 0; wd < kWords; ++wd) a[wd] |= w4[wd];
                        }
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) {
                            a[wd] = __reduce_or_sync(0xFFFFFFFFu, a[wd]);
                            if (lane == o) acc[wd] = a[wd];
                        }
                    }
                    uint32_t fresh[kWords], any_fresh = 0, still = 0;
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) { fresh[wd] = acc[wd] & unseen[wd]; any_fresh |= fresh[wd]; still |= unseen[wd] & ~fresh[wd]; }
                    if (any_fresh) WT::stcg(next_of(x, par_n), WT::pack(fresh));     // one writer per vertex
                    my_open += still != 0 ? 1u : 0u;
                    const bool enqueue = any_fresh != 0;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, enqueue);
                    if (bal) {
                        uint32_t qb = 0;
                        const int leader = __ffs(bal) - 1;
                        if (lane == leader) qb = atomicAdd(tail_p, (uint32_t)__popc(bal));
                        qb = __shfl_sync(0xFFFFFFFFu, qb, leader);
                        if (enqueue) {
                            const uint32_t pos = qb + __popc(bal & lane_lt);
                            if (pos < qcap) queue_v[pos] = x;
                            else *(volatile uint32_t *)fail_p = 1;
                        }
                    }
                }
                my_open = __reduce_add_sync(0xFFFFFFFFu, my_open);
                if (lane == 0 && my_open) atomicAdd(open_p, my_open);
            }
            prev_pull = pull;
            if (tid_t == 0) dir_levels += pull ? (1ull << 32) : 1ull;
            ++level;
        }
        // here: levels 0..level-1 are non-empty (if !too_deep), lvl[l] = start of level l, lvl[level] = end of the last
        if constexpr (kVirtual) for (uint32_t L = tid_t; L <= level; L += threads_t) bcnt[L] = 0;
        team_sync();
        const bool failed = *(volatile uint32_t *)fail_p != 0;
        if (too_deep || failed) {
            if (tid_t == 0) p.batch_status[batch] = failed ? kQueueOverflow : kTooDeep;
            team_sync();                              // (the reset of the next batch clears every record)
            continue;
        }
        const uint32_t depth = level - 1;
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_fwd += t - t_phase; t_phase = t; }
        if (depth > my_depth) my_depth = depth;
        if (p.forward_only) {
            if (tid_t == 0) { atomicAdd(&s_cnt[2], (unsigned long long)qend); p.batch_status[batch] = kDone; }
            team_sync();
            continue;
        }

        // ---- backward: levels depth .. 1 ----------------------------------------------------------------------
        // prepare(L): every entry of level L gets the position of its run in qc (prefix sums per warp, one atomic on
        // the team's counter) and publishes {mask, position, tag = L + 1} in slot L of its vertex's level record.
        // statistics (pairs reached, adjacency entries in reach, queue entries): summed per tile of 32 entries
        auto count_tile = [&](uint32_t c, uint32_t deg, bool have) {
            const uint32_t csum = __reduce_add_sync(0xFFFFFFFFu, c);
            unsigned long long e = (unsigned long long)c * deg;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
            const uint32_t nv = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, have));
            if (lane == 0) {
                atomicAdd(&s_cnt[0], (unsigned long long)csum);
                atomicAdd(&s_cnt[1], e);
                atomicAdd(&s_cnt[2], (unsigned long long)nv);
            }
        };
        auto prepare = [&](uint32_t L) {
            const uint32_t lb = lvl[L], le = lvl[L + 1];
            const uint32_t phys = ((L & 1u) << 1) | ((L >> 1) & 1u);
            for (uint32_t base = lb + warp_t * 32; base < le; base += warps_t * 32) {
                const uint32_t idx = base + lane;
                uint32_t m[kWords], c = 0, u = 0;
#pragma unroll
                for (int wd = 0; wd < kWords; ++wd) m[wd] = 0;
                if (idx < le) {
                    u = queue_v[idx];
                    WT::unpack(queue_m[idx], m);
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) c += __popc(m[wd]);
                }
                uint32_t pos = 0;
                if (L >= 1) {                           // the sources themselves (level 0) store no q
                    const uint32_t incl = warp_inclusive_scan(c, lane);
                    uint32_t b = 0;
                    if (lane == 31) b = atomicAdd(pos_p, incl);
                    b = __shfl_sync(0xFFFFFFFFu, b, 31);
                    pos = b + incl - c;
                }
                if (idx < le) {
                    queue_p[idx] = pos;
                    uint4 *dst = reinterpret_cast<uint4 *>(lr + (size_t)u * 4 * kSlotB + phys * kSlotB);
                    if constexpr (kW == 256) {
                        __stcg(dst, make_uint4(m[0], m[1], m[2], m[3]));
                        __stcg(dst + 1, make_uint4(m[4], m[5], m[6], m[7]));
                        __stcg(dst + 2, make_uint4(pos, L + 1u, 0u, 0u));
                    } else if constexpr (kW == 128) {
                        __stcg(dst, make_uint4(m[0], m[1], m[2], m[3]));
                        __stcg(dst + 1, make_uint4(pos, L + 1u, 0u, 0u));
                    } else
                        __stcg(dst, make_uint4(m[0], m[1], pos, L + 1u));
                }
                if constexpr (kVirtual) {
                    // the level's work list: one item per entry, two (ranks 0..127, ranks 128..) if it carries more than 128 sources
                    if (L >= 1) {
                        const uint32_t items = idx < le ? (c > 128u ? 2u : 1u) : 0u;
                        const uint32_t incl = warp_inclusive_scan(items, lane);
                        uint32_t b = 0;
                        if (lane == 31 && incl) b = atomicAdd(&bcnt[L], incl);
                        b = __shfl_sync(0xFFFFFFFFu, b, 31);
                        const uint32_t at = 2u * lb + b + incl - items;
                        if (items >= 1u) bq[at] = idx;
                        if (items == 2u) bq[at + 1] = idx | 0x80000000u;
                    }
                }
                if (L == 0) {                           // statistics of the level the sweep below never visits
                    uint32_t deg0 = 0;
                    if (idx < le) deg0 = __ldg(p.row_ptr + u + 1) - __ldg(p.row_ptr + u);
                    count_tile(c, deg0, idx < le);
                }
            }
        };
        prepare(depth);
        for (uint32_t l = depth; l >= 1; --l) {
            prepare(l - 1);
            team_sync();                              // level l+1 fully processed (its q values stored), levels l and l-1 prepared
            uint32_t lb = lvl[l], le = lvl[l + 1];
            if constexpr (kVirtual) { le = 2u * lb + __ldcg(&bcnt[l]); lb = 2u * lb; }      // items of the level's work list
            const uint32_t pair_off = ((l + 1u) & 1u) * 2u * kSlotB;       // block of the slots of levels l+1 and l-1
            const uint32_t sub_s = ((l + 1u) >> 1) & 1u;                   // which of the two is level l+1's
            const uint32_t tag_s = l + 2u, tag_p = l;
            for (uint32_t tile = lb + warp_t * 32; tile < le; tile += warps_t * 32) {
                const uint32_t idx = tile + lane;
                __syncwarp();                                       // the previous tile's staged state is no longer read
                uint32_t total, my_qi = 0;                          // my_qi: queue index of this lane's item (bit 31: second half of its entry)
                {
                    uint32_t rs = 0, deg = 0, c = 0;
                    bool counted = idx < le;
                    if (idx < le) {
                        uint32_t qi = idx, half = 0;
                        if constexpr (kVirtual) { const uint32_t it = __ldcg(&bq[idx]); qi = it & 0x7FFFFFFFu; half = it >> 31; }
                        my_qi = qi | (half << 31);
                        const uint32_t w = queue_v[qi];
                        uint32_t m[kWords];
                        WT::unpack(queue_m[qi], m);
                        uint32_t pos = queue_p[qi];
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) c += __popc(m[wd]);
                        rs = __ldg(p.row_ptr + w);
                        deg = __ldg(p.row_ptr + w + 1) - rs;
                        if constexpr (kVirtual) {
                            if (c > 128u) {                         // this item's half of the entry: the sources of rank < 128, or the rest
                                uint32_t cum = 0;
#pragma unroll
                                for (int wd = 0; wd < kWords; ++wd) {
                                    const uint32_t pc = __popc(m[wd]);
                                    uint32_t rest = m[wd];          // what belongs to the second half
                                    if (cum + pc <= 128u) rest = 0u;
                                    else if (cum < 128u) for (uint32_t k = 128u - cum; k; --k) rest &= rest - 1u;   // drop the k lowest set bits
                                    m[wd] = half ? rest : m[wd] ^ rest;
                                    cum += pc;
                                }
                                if (half) { pos += 128u; counted = false; }
                            }
                        }
                        *reinterpret_cast<Vec *>(&my_tile_m[lane][0]) = WT::pack(m);
                        s_tile_p[warp][lane] = pos;
                    }
                    count_tile(counted ? c : 0u, deg, counted);
                    const uint32_t incl = warp_inclusive_scan(deg, lane);
                    total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    s_tile_base[warp][lane] = rs - (incl - deg);
                    s_tile_end[warp][lane] = incl;
                }
                __syncwarp();
                // per-entry state of this lane: up to kWords sources (ranks lane, lane + 32, ...)
                int cur = -1;
                uint32_t cur_end = 0, c_cur = 0, pos_cur = 0;
                uint32_t cell[kLaneSlots], sel[kLaneSlots], npred[kLaneSlots];     // byte offset of the source's cell in a cooked row, its bit, predecessors so far
                double acc[kLaneSlots];
#pragma unroll
                for (int h = 0; h < kLaneSlots; ++h) { cell[h] = 0; sel[h] = 0; npred[h] = 0; acc[h] = 0.0; }

                auto begin_entry = [&]() {
                    uint32_t mc[kWords], pre[kWords + 1];
                    pre[0] = 0;
                    WT::unpack(*reinterpret_cast<const Vec *>(&my_tile_m[cur][0]), mc);      // broadcast reads
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) pre[wd + 1] = pre[wd] + __popc(mc[wd]);
                    c_cur = pre[kWords];
                    pos_cur = s_tile_p[warp][cur];
                    __syncwarp();                                   // the previous entry's list has been read by every lane
                    {
                        const unsigned kl = (unsigned)__cvta_generic_to_shared(&s_klist[warp][0]);
#pragma unroll
                        for (int wd = 0; wd < kWords; ++wd) {       // predicated byte stores, no branches
                            const unsigned at = kl + pre[wd] + __popc(mc[wd] & lane_lt);
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t}"
                                         ::"r"(at), "r"((uint32_t)(wd * 32 + lane)), "r"((mc[wd] >> lane) & 1u) : "memory");
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < kLaneSlots; ++h) {
                        const bool on = (uint32_t)lane + 32u * h < c_cur;
                        const uint32_t ks = on ? (uint32_t)s_klist[warp][lane + 32 * h] : 0u;
                        cell[h] = (ks >> 5) * 16u;
                        sel[h] = on ? 1u << (ks & 31u) : 0u;
                        acc[h] = 0.0; npred[h] = 0;
                    }
                };
                auto finish_entry = [&]() {
                    double contrib = 0.0;
#pragma unroll
                    for (int h = 0; h < kLaneSlots; ++h) {
                        if ((uint32_t)lane + 32u * h < c_cur) {
                            const double bk = 1.0 + acc[h];
                            __stcg(&qc[(size_t)pos_cur + lane + 32 * h], bk / (double)npred[h]);    // dense, coalesced
                            contrib += p.include_endpoints ? bk : bk - 1.0;
                        }
                    }
                    const double tot = warp_sum(contrib);
                    if (lane == 0) *reinterpret_cast<double *>(&my_tile_m[cur][0]) = tot;      // the mask was consumed by begin_entry
                };
                // kRows rows of the current entry starting at row r of the cooked chunk in `rows`, kSlots sources per
                // lane.  Straight-line: every q load of the group is issued (predicated) before the first is consumed.
                auto consume = [&](auto rows_tag, auto slots_tag, const unsigned char *rows, uint32_t r, uint32_t nrows) {
                    constexpr int kRows = decltype(rows_tag)::value, kSlots = decltype(slots_tag)::value;
                    double qv[kRows][kSlots];
                    const unsigned char *rp = rows + r * kRowS;
#pragma unroll
                    for (int jj = 0; jj < kRows; ++jj) {
                        const bool valid = (uint32_t)jj < nrows;    // warp-uniform
#pragma unroll
                        for (int h = 0; h < kSlots; ++h) {
                            const uint4 cl = *reinterpret_cast<const uint4 *>(rp + jj * kRowS + cell[h]);   // {succ word, pred word, base, -}
                            const bool succ = valid && (cl.x & sel[h]) != 0u;
                            const uint32_t at = cl.z + __popc(cl.x & (sel[h] - 1u));
                            qv[jj][h] = succ ? __ldcg(qc_hot + at) : 0.0;
                            if (valid && (cl.y & sel[h]) != 0u) npred[h] += 1u;
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < kRows; ++jj)              // adjacency (CSR) order: fixes the rounding
#pragma unroll
                        for (int h = 0; h < kSlots; ++h) acc[h] += qv[jj][h];
                };

                auto edge_of = [&](uint32_t t0) -> uint32_t {     // neighbour id of flattened entry t0 + lane
                    const uint32_t t = t0 + lane;
                    if (t >= total || (uint32_t)lane >= kChunk) return kWideInvalid;
                    int o = 0;                                    // first entry whose rows end beyond t (zero-row entries are skipped)
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1)
                        if (s_tile_end[warp][o + step - 1] <= t) o += step;
                    return __ldg(p.col_idx + (uint32_t)(s_tile_base[warp][o] + t));     // (32-bit wrap-around: the base may be 'negative')
                };
                auto stage_rows = [&](int buf, uint32_t xs) {     // level-record blocks of the chunk whose neighbour ids are `xs` (lane r: row r)
#pragma unroll
                    for (int it = 0; it < ((int)kChunk * kPieces + 31) / 32; ++it) {      // 32 / kPieces rows per step
                        const int r = it * (32 / kPieces) + lane / kPieces;
                        const int part = lane % kPieces;
                        const uint32_t x = __shfl_sync(0xFFFFFFFFu, xs, r);
                        if (x != kWideInvalid) {
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(my_rows + ((size_t)buf * kChunk + r) * kRowS + part * 16);
                            const unsigned char *src = lr + (size_t)x * 4 * kSlotB + pair_off + part * 16;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                // bulk form: lane r requests row r as one 128 B copy; the warp's mbarrier counts the bytes
                auto stage_rows_bulk = [&](uint32_t xs, uint32_t cnt) {
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[warp]);
                    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * (uint32_t)kRowB) : "memory");
                    __syncwarp();
                    if (xs != kWideInvalid) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(my_rows + (size_t)lane * kRowS);
                        const unsigned char *src = lr + (size_t)xs * 4 * kSlotB + pair_off;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst), "l"(src), "r"((uint32_t)kRowB), "r"(bar) : "memory");
                    }
                };
                auto wait_rows_bulk = [&]() {
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[warp]);
                    const uint32_t phase = s_mphase[warp];
                    uint32_t done;
                    do {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
                    } while (!done);
                    __syncwarp();                                   // every lane has read the parity
                    if (lane == 0) s_mphase[warp] = phase ^ 1u;
                };
                // lane r turns the landed block of row r into kWords cells {succ mask word, pred mask word, index of the
                // word's first value in the neighbour's run, -}; slots with another level's tag count as empty
                auto cook = [&](int buf, bool row_valid) {
                    if ((uint32_t)lane >= kChunk) return;           // (no such row: the address would belong to the next warp)
                    uint4 *rp = reinterpret_cast<uint4 *>(my_rows + ((size_t)buf * kChunk + lane) * kRowS);
                    uint32_t ms[kWords], mp[kWords], pos_s, ts, tp;
                    if constexpr (kW == 256) {
                        const uint4 a0 = rp[0], a1 = rp[1], a2 = rp[2], b0 = rp[4], b1 = rp[5], b2 = rp[6];
                        const uint4 s0 = sub_s ? b0 : a0, s1 = sub_s ? b1 : a1, s2 = sub_s ? b2 : a2;
                        const uint4 p0 = sub_s ? a0 : b0, p1 = sub_s ? a1 : b1, p2 = sub_s ? a2 : b2;
                        ms[0] = s0.x; ms[1] = s0.y; ms[2] = s0.z; ms[3] = s0.w; ms[4] = s1.x; ms[5] = s1.y; ms[6] = s1.z; ms[7] = s1.w; pos_s = s2.x; ts = s2.y;
                        mp[0] = p0.x; mp[1] = p0.y; mp[2] = p0.z; mp[3] = p0.w; mp[4] = p1.x; mp[5] = p1.y; mp[6] = p1.z; mp[7] = p1.w; tp = p2.y;
                    } else if constexpr (kW == 128) {
                        const uint4 a0 = rp[0], a1 = rp[1], b0 = rp[2], b1 = rp[3];
                        const uint4 s0 = sub_s ? b0 : a0, s1 = sub_s ? b1 : a1, p0 = sub_s ? a0 : b0, p1 = sub_s ? a1 : b1;
                        ms[0] = s0.x; ms[1] = s0.y; ms[2] = s0.z; ms[3] = s0.w; pos_s = s1.x; ts = s1.y;
                        mp[0] = p0.x; mp[1] = p0.y; mp[2] = p0.z; mp[3] = p0.w; tp = p1.y;
                    } else {
                        const uint4 a = rp[0], b = rp[1];
                        const uint4 s = sub_s ? b : a, q = sub_s ? a : b;
                        ms[0] = s.x; ms[1] = s.y; pos_s = s.z; ts = s.w;
                        mp[0] = q.x; mp[1] = q.y; tp = q.w;
                    }
                    const bool s_ok = row_valid && ts == tag_s, p_ok = row_valid && tp == tag_p;
                    uint32_t at = pos_s;
#pragma unroll
                    for (int wd = 0; wd < kWords; ++wd) {
                        const uint32_t a = s_ok ? ms[wd] : 0u, b = p_ok ? mp[wd] : 0u;
                        rp[wd] = make_uint4(a, b, at, 0u);
                        at += __popc(a);
                    }
                };

                // Two row buffers: chunk sc+1's blocks are in flight while chunk sc is consumed.  One buffer (256 sources per
                // batch: a row is 128 B): the chunk's blocks are requested and waited for in turn; the neighbour ids are
                // loaded two chunks ahead either way.
                const uint32_t n_chunks = (total + kChunk - 1) / kChunk;
                uint32_t x_cur = edge_of(0);                               // neighbour ids of chunk sc
                uint32_t x_next = n_chunks > 1 ? edge_of(kChunk) : kWideInvalid;   // ... and of chunk sc+1
                if constexpr (kBufs == 2) stage_rows(0, x_cur);
                for (uint32_t sc = 0; sc < n_chunks; ++sc) {
                    const int buf = kBufs == 2 ? (int)(sc & 1) : 0;
                    const uint32_t t0 = sc * kChunk;
                    const uint32_t cnt = min(kChunk, total - t0);
                    if constexpr (kBulk) { stage_rows_bulk(x_cur, cnt); wait_rows_bulk(); }
                    else {
                        if constexpr (kBufs == 1) stage_rows(0, x_cur);
                        asm volatile("cp.async.wait_group 0;" ::: "memory");   // chunk sc's blocks have landed (this lane's copies)
                    }
                    __syncwarp();                                          // ... and every other lane's
                    if constexpr (kBufs == 2) stage_rows(buf ^ 1, x_next); // chunk sc+1 (nothing past the end); buffer drained at sc-1
                    x_cur = x_next;
                    x_next = sc + 2 < n_chunks ? edge_of(t0 + 2 * kChunk) : kWideInvalid;
                    cook(buf, (uint32_t)lane < cnt);
                    __syncwarp();
                    const unsigned char *rows = my_rows + (size_t)buf * kChunk * kRowS;
                    for (uint32_t r = 0; r < cnt;) {
                        const uint32_t tt = t0 + r;
                        if (tt >= cur_end) {                      // warp-uniform: the next entry's rows start
                            if (cur >= 0) finish_entry();
                            do { ++cur; cur_end = s_tile_end[warp][cur]; } while (tt >= cur_end);
                            begin_entry();
                        }
                        const uint32_t left = min(cur_end - tt, cnt - r);      // rows of this entry inside this chunk
                        if (kLaneSlots == 4 && c_cur > 96) {
                            const uint32_t nrows = min(left, 2u);
                            consume(std::integral_constant<int, 2>{}, std::integral_constant<int, kLaneSlots>{}, rows, r, nrows);
                            r += nrows;
                        } else if (kLaneSlots == 4 && c_cur > 64) {
                            const uint32_t nrows = min(left, 2u);
                            consume(std::integral_constant<int, 2>{}, std::integral_constant<int, 3>{}, rows, r, nrows);
                            r += nrows;
                        } else if (c_cur > 32) {
                            const uint32_t nrows = min(left, 4u);
                            consume(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{}, rows, r, nrows);
                            r += nrows;
                        } else if (left > 4u) {
                            const uint32_t nrows = min(left, 8u);
                            consume(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{}, rows, r, nrows);
                            r += nrows;
                        } else {                                  // a short tail: the half-size group wastes fewer predicated-off rows
                            consume(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{}, rows, r, left);
                            r += left;
                        }
                    }
                    __syncwarp();                                 // buffer `buf` may be refilled next iteration
                    if constexpr (kBulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cooked (generic) writes before the next bulk copies
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                if (cur >= 0) finish_entry();
                __syncwarp();
                if (idx < le) {                                     // one writer per vertex, level and half
                    double *b = bslot;
                    if constexpr (kVirtual) { if (my_qi >> 31) b = bslot + p.bslot_hi_off; }
                    // (an entry of level >= 1 has at least its predecessor's row; one without rows was never begun and still holds its mask)
                    if (s_tile_end[warp][lane] != (lane ? s_tile_end[warp][lane - 1] : 0u))
                        b[queue_v[my_qi & 0x7FFFFFFFu]] += *reinterpret_cast<const double *>(&my_tile_m[lane][0]);
                }
            }
        }
        if (threadIdx.x == 0) { if (team_rank == 0) p.batch_status[batch] = kDone; cyc_bwd += clock64() - t_phase; }
        team_sync();
    }

    // ---- flush counters ---------------------------------------------------------------------------------------
    team_sync();                                      // rank 0's shared memory stays alive until every CTA of the team is done
    if (threadIdx.x < 3) atomicAdd(&p.counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x == 0) {
        atomicMax(&p.counters[3], (unsigned long long)my_depth);
        atomicAdd(&p.counters[4], (unsigned long long)cyc_reset);
        atomicAdd(&p.counters[5], (unsigned long long)cyc_fwd);
        atomicAdd(&p.counters[6], (unsigned long long)cyc_bwd);
        if (dir_levels) atomicAdd(&p.counters[7], dir_levels);
    }
}

}  // namespace nec
