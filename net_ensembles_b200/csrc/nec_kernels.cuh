// CUDA kernels of the vertex_load hot path (sm_100a).
//
// What is computed (reference: src/generic_graph/generic_graph.rs:1310-1385): for every
// source s a BFS gives d_s(.); with P_s(w) = {u in N(w): d_s(u) = d_s(w)-1} the reverse
// sweep yields  bk_s(w) = 1 + sum_{x: w in P_s(x)} bk_s(x)/|P_s(x)|  and the result is
// b[v] = sum_{s != v, s reaches v} (bk_s(v) - [!include_endpoints]).
//
// How it is laid out for the GPU — "batched multi-source, source-minor":
//   * 32 sources form a BATCH; lane k of every warp owns source k of the batch.
//   * per-batch state is stored source-minor: dist[v][32] (u8, 32 B = one sector per vertex
//     row; u32 for graphs deeper than 253 levels) and q[v][32] (f64, 256 B per row), so a
//     neighbour gather by a warp is one fully used sector / two fully used lines, and the
//     CSR row of a vertex is read once for 32 sources.
//   * one CTA owns one batch SLOT (its arena of dist/q/queues) and walks the batch
//     level-synchronously with __syncthreads only; CTAs are persistent and take batches
//     round-robin (static: batch -> slot mapping and therefore every sum is reproducible).
//   * forward pass (push): frontier queue per level; an entry is (vertex, mask of lanes whose
//     BFS has the vertex on this level).  A warp expands one entry: for each neighbour row it
//     marks the lanes that see it for the first time (dist = level+1, plain byte stores, all
//     racing writers store the same value) and, warp-ballot, ORs the lane mask into
//     nextmask[x]; the warp that turns nextmask[x] from 0 to non-zero appends x to the next
//     level's queue (shared-memory tail counter).
//   * backward pass (pull, atomic-free, deterministic): levels deepest -> 1, same queues.  The
//     warp owning (w, level l) scans w's row once: lanes with dist[x]==l+1 pull q[x], lanes
//     count dist[x]==l-1 as predecessors; bk = 1 + sum, q[w] = bk / npred, and the lane sum of
//     bk is added to this slot's private b[] (one writer per vertex per level => no atomics).
//     Predecessor lists are never materialised; sigma never enters the value.
//   * slots' private b[] are summed in slot order by reduce_slots_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nec {

constexpr int kLanes = 32;            // sources per batch
constexpr int kEngineThreads = 512;   // 16 warps per CTA
constexpr int kEngineMinBlocks = 2;   // CTAs per SM the register budget is sized for
constexpr int kUnroll = 4;            // neighbour rows in flight per warp step

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
    // per-slot arenas (slot = blockIdx.x)
    uint8_t *dist_base;   size_t dist_stride;   // bytes per slot
    double *q_base;       size_t q_stride;      // doubles per slot
    uint32_t *mask_base;  size_t mask_stride;   // u32 per slot (2n: two level parities)
    double *bslot_base;   size_t bslot_stride;  // doubles per slot (n)
    uint2 *queue_base;    size_t queue_stride;  // entries per slot (= capacity)
    uint32_t *lvl_base;   size_t lvl_stride;    // u32 per slot (max_levels + 2)
    uint32_t max_levels;                        // deepest level index the slot can record
    // outputs
    uint8_t *batch_status;                      // per batch id
    unsigned long long *counters;               // [0] reached pairs [1] edge pairs [2] visits [3] max depth
    // optional per-(vertex,lane) debug planes of slot 0 (nec_bfs_debug)
    uint32_t *dbg_npred;
    double *dbg_bk;
};

template <typename DistT> struct DistTraits;
template <> struct DistTraits<uint8_t>  { static constexpr uint32_t kInf = 0xFFu;        static constexpr uint32_t kMaxLevel = 253u; };
template <> struct DistTraits<uint32_t> { static constexpr uint32_t kInf = 0xFFFFFFFFu;  static constexpr uint32_t kMaxLevel = 0xFFFFFFF0u; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// The engine.  One CTA = one slot; loops over its batches.
template <typename DistT, bool kDebug>
__global__ void __launch_bounds__(kEngineThreads, kEngineMinBlocks)
vertex_load_engine(const EngineParams p) {
    constexpr uint32_t kInf = DistTraits<DistT>::kInf;
    constexpr int kWarps = kEngineThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t n = p.n;

    DistT *__restrict__ dist = reinterpret_cast<DistT *>(p.dist_base + (size_t)blockIdx.x * p.dist_stride);
    double *__restrict__ q = p.q_base + (size_t)blockIdx.x * p.q_stride;
    uint32_t *mask_even = p.mask_base + (size_t)blockIdx.x * p.mask_stride;
    uint32_t *mask_odd = mask_even + n;
    double *__restrict__ bslot = p.bslot_base + (size_t)blockIdx.x * p.bslot_stride;
    uint2 *__restrict__ queue = p.queue_base + (size_t)blockIdx.x * p.queue_stride;
    uint32_t *__restrict__ lvl = p.lvl_base + (size_t)blockIdx.x * p.lvl_stride;
    const uint32_t qcap = (uint32_t)p.queue_stride;
    const uint32_t level_limit = p.max_levels < DistTraits<DistT>::kMaxLevel ? p.max_levels : DistTraits<DistT>::kMaxLevel;

    __shared__ uint32_t s_tail;
    __shared__ uint32_t s_fail;
    __shared__ unsigned long long s_cnt[3];

    unsigned long long my_reached = 0, my_edges = 0, my_visits = 0;  // lane 0 of each warp
    uint32_t my_depth = 0;

    for (uint32_t work = blockIdx.x; work < p.n_work; work += gridDim.x) {
        const uint32_t batch = p.work_list ? p.work_list[work] : work;
        unsigned long long b_reached = 0, b_edges = 0, b_visits = 0;  // counted only if the batch completes

        // ---- reset this slot's distances (streaming 16 B stores) ----------------------
        {
            const size_t n16 = ((size_t)n * kLanes * sizeof(DistT)) / 16;
            uint4 *d16 = reinterpret_cast<uint4 *>(dist);
            const uint4 inf4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            for (size_t i = threadIdx.x; i < n16; i += kEngineThreads) d16[i] = inf4;
        }
        if (threadIdx.x == 0) { s_tail = 0; s_fail = 0; }
        __syncthreads();

        // ---- seed level 0: lane k <-> source k of the batch ---------------------------
        if (warp == 0) {
            const uint32_t sidx = batch * kLanes + lane;
            if (sidx < p.n_sources) {
                const uint32_t s = p.sources[sidx];
                dist[(size_t)s * kLanes + lane] = (DistT)0;
                const uint32_t old = atomicOr(&mask_even[s], 1u << lane);
                if (old == 0) queue[atomicAdd(&s_tail, 1u)].x = s;
            }
        }
        __syncthreads();

        // ---- forward: level-synchronous push --------------------------------------------
        uint32_t level = 0, qbeg = 0, qend = 0;
        bool too_deep = false;
        for (;;) {
            __syncthreads();                         // previous level fully expanded
            const uint32_t tail_now = s_tail < qcap ? s_tail : qcap;
            __syncthreads();                         // every warp sampled the tail before it moves again
            qbeg = qend;
            qend = tail_now;
            if (threadIdx.x == 0) lvl[level] = qbeg;
            if (qend == qbeg) break;                 // level `level` is empty: done
            if (level >= level_limit) { too_deep = true; break; }
            uint32_t *mask_cur = (level & 1) ? mask_odd : mask_even;
            uint32_t *mask_nxt = (level & 1) ? mask_even : mask_odd;
            const DistT next_d = (DistT)(level + 1);

            for (uint32_t idx = qbeg + warp; idx < qend; idx += kWarps) {
                uint32_t u = 0, m = 0;
                if (lane == 0) {
                    u = queue[idx].x;
                    m = mask_cur[u];
                    mask_cur[u] = 0;                 // self-cleaning: parity is reused at level+2
                    queue[idx].y = m;                // the backward pass reads (vertex, mask)
                }
                u = __shfl_sync(0xFFFFFFFFu, u, 0);
                m = __shfl_sync(0xFFFFFFFFu, m, 0);
                const bool active = (m >> lane) & 1u;
                const uint32_t rs = __ldg(p.row_ptr + u), re = __ldg(p.row_ptr + u + 1);
                if (lane == 0) {
                    const uint32_t pc = __popc(m);
                    b_reached += pc;
                    b_edges += (unsigned long long)pc * (re - rs);
                    b_visits += 1;
                }
                for (uint32_t e = rs; e < re; e += 32) {
                    const uint32_t c = min(32u, re - e);
                    const uint32_t xl = lane < c ? __ldg(p.col_idx + e + lane) : 0u;
                    for (uint32_t j0 = 0; j0 < c; j0 += kUnroll) {
                        uint32_t x[kUnroll];
                        uint32_t dx[kUnroll];
#pragma unroll
                        for (int jj = 0; jj < kUnroll; ++jj) {
                            x[jj] = __shfl_sync(0xFFFFFFFFu, xl, (j0 + jj) & 31);
                            dx[jj] = (j0 + jj < c) ? (uint32_t)dist[(size_t)x[jj] * kLanes + lane] : 0u;
                        }
#pragma unroll
                        for (int jj = 0; jj < kUnroll; ++jj) {
                            const bool fresh = active && dx[jj] == kInf;
                            if (fresh) dist[(size_t)x[jj] * kLanes + lane] = next_d;
                            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, fresh);
                            if (bal != 0 && lane == 0) {
                                const uint32_t old = atomicOr(&mask_nxt[x[jj]], bal);
                                if (old == 0) {
                                    const uint32_t pos = atomicAdd(&s_tail, 1u);
                                    if (pos < qcap) queue[pos].x = x[jj];
                                    else s_fail = 1;
                                }
                            }
                        }
                    }
                }
            }
            ++level;
        }
        // here: levels 0..level-1 are non-empty (if !too_deep), lvl[l] = start of level l,
        // lvl[level] = end of the last level.
        __syncthreads();
        const bool failed = s_fail != 0;

        if (too_deep || failed) {
            // leave the slot clean for the next batch: drop pending mask bits (rare path)
            for (uint32_t v = threadIdx.x; v < n; v += kEngineThreads) { mask_even[v] = 0; mask_odd[v] = 0; }
            if (threadIdx.x == 0) p.batch_status[batch] = failed ? kQueueOverflow : kTooDeep;
            __syncthreads();
            continue;
        }
        my_reached += b_reached; my_edges += b_edges; my_visits += b_visits;
        const uint32_t depth = level - 1;             // deepest non-empty level
        if (depth > my_depth) my_depth = depth;

        // ---- backward: levels depth .. 1, pull from successors --------------------------
        for (uint32_t l = depth; l >= 1; --l) {
            const uint32_t lb = lvl[l], le = lvl[l + 1];
            const uint32_t d_succ = l + 1, d_pred = l - 1;
            for (uint32_t idx = lb + warp; idx < le; idx += kWarps) {
                const uint2 ent = queue[idx];
                const uint32_t w = ent.x;
                const bool active = (ent.y >> lane) & 1u;
                const uint32_t rs = __ldg(p.row_ptr + w), re = __ldg(p.row_ptr + w + 1);
                double acc = 0.0;
                uint32_t npred = 0;
                for (uint32_t e = rs; e < re; e += 32) {
                    const uint32_t c = min(32u, re - e);
                    const uint32_t xl = lane < c ? __ldg(p.col_idx + e + lane) : 0u;
                    for (uint32_t j0 = 0; j0 < c; j0 += kUnroll) {
                        uint32_t x[kUnroll];
                        uint32_t dx[kUnroll];
                        double qv[kUnroll];
#pragma unroll
                        for (int jj = 0; jj < kUnroll; ++jj) {
                            x[jj] = __shfl_sync(0xFFFFFFFFu, xl, (j0 + jj) & 31);
                            dx[jj] = (j0 + jj < c) ? (uint32_t)dist[(size_t)x[jj] * kLanes + lane] : kInf;
                        }
#pragma unroll
                        for (int jj = 0; jj < kUnroll; ++jj)
                            qv[jj] = (active && dx[jj] == d_succ) ? q[(size_t)x[jj] * kLanes + lane] : 0.0;
#pragma unroll
                        for (int jj = 0; jj < kUnroll; ++jj) {   // CSR order: fixes the rounding
                            acc += qv[jj];
                            npred += (dx[jj] == d_pred) ? 1u : 0u;
                        }
                    }
                }
                const double bk = 1.0 + acc;
                double contrib = 0.0;
                if (active) {
                    q[(size_t)w * kLanes + lane] = bk / (double)npred;
                    contrib = p.include_endpoints ? bk : bk - 1.0;
                    if (kDebug) {
                        p.dbg_npred[(size_t)w * kLanes + lane] = npred;
                        p.dbg_bk[(size_t)w * kLanes + lane] = bk;
                    }
                }
                const double total = warp_sum(contrib);
                if (lane == 0) bslot[w] += total;
            }
            __syncthreads();
        }
        if (kDebug) {
            // the sources themselves (level 0) get bk but no load contribution
            const uint32_t lb = lvl[0], le = lvl[1];
            for (uint32_t idx = lb + warp; idx < le; idx += kWarps) {
                const uint2 ent = queue[idx];
                const bool active = (ent.y >> lane) & 1u;
                const uint32_t rs = p.row_ptr[ent.x], re = p.row_ptr[ent.x + 1];
                double acc = 0.0;
                for (uint32_t e = rs; e < re; ++e) {
                    const uint32_t x = p.col_idx[e];
                    if (active && (uint32_t)dist[(size_t)x * kLanes + lane] == 1u) acc += q[(size_t)x * kLanes + lane];
                }
                if (active) {
                    p.dbg_npred[(size_t)ent.x * kLanes + lane] = 0;
                    p.dbg_bk[(size_t)ent.x * kLanes + lane] = 1.0 + acc;
                }
            }
        }
        if (threadIdx.x == 0) p.batch_status[batch] = kDone;
        __syncthreads();
    }

    // ---- flush counters -------------------------------------------------------------------
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
        atomicAdd(&s_cnt[0], my_reached);
        atomicAdd(&s_cnt[1], my_edges);
        atomicAdd(&s_cnt[2], my_visits);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&p.counters[threadIdx.x], s_cnt[threadIdx.x]);
    if (threadIdx.x == 0) atomicMax(&p.counters[3], (unsigned long long)my_depth);
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

// a[v] += b[v]
__global__ void add_into_kernel(double *__restrict__ a, const double *__restrict__ b, uint32_t n) {
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) a[v] += b[v];
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

// Shortest-path counts from given distances, one CTA, level by level (debug only; sigma is
// not part of the load value).  sigma[v] = sum of sigma over neighbours one level closer.
__global__ void sigma_debug_kernel(uint32_t n, const uint32_t *__restrict__ row_ptr,
                                   const uint32_t *__restrict__ col_idx, const uint32_t *__restrict__ dist,
                                   uint32_t max_depth, double *sigma) {
    for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) sigma[v] = dist[v] == 0 ? 1.0 : 0.0;
    __syncthreads();
    for (uint32_t l = 1; l <= max_depth; ++l) {
        for (uint32_t v = threadIdx.x; v < n; v += blockDim.x) {
            if (dist[v] != l) continue;
            double acc = 0.0;
            for (uint32_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                const uint32_t x = col_idx[e];
                if (dist[x] == l - 1) acc += sigma[x];
            }
            sigma[v] = acc;
        }
        __syncthreads();
    }
}

}  // namespace nec
