// "Mid" engine: one source per CTA with the per-source integer state in shared memory.
// Used when a whole distance vector fits on chip (n <= kMidMaxN): the regime of BASELINE
// config 2 (small-world N = 2^16), where 32-source batches are incoherent (a vertex sits
// on ~7 different levels within a batch) and the batched engine wastes 6/7 of its lanes.
//
// Same arithmetic as everywhere else (reference: generic_graph.rs:1310-1385): BFS distances,
// then deepest level first  bk(w) = 1 + sum_{x in N(w), d(x)=d(w)+1} q(x),  q(w) = bk(w)/npred(w),
// sums taken in adjacency (CSR) order by ONE thread per vertex, so the result is
// reproducible run to run.  No predecessor lists, no sigma, no atomics on floating point.
//
// Memory placement per CTA (= per in-flight source):
//   shared : dist[n] u8, level offsets, per-thread slab staging slots (16 B chunks, cp.async targets)
//   global : q[n] f64      — the randomly read global array; it is written on level l+1 and
//                            read on level l, so only ~two levels' worth of sectors per CTA has
//                            to survive in the 126 MB L2
//            order[n] u32  — BFS order = the level queues; written/read as a stream
//            pos4[n] 16 B  — per queue position: the vertex's first four BFS successors with its successor
//                            count, predecessor count and "row needed" flag packed into the spare high
//                            bits (ids < 2^18); one coalesced 16 B store / load; streamed
//            bslot[n] f64  — this CTA's private partial load, updated by a coalesced sweep
//                            after each source (streams through HBM; evict-first hints)
// Forward (thread per frontier vertex): one slab row (cp.async-staged one step ahead) gives the
// degree and the first W-1 neighbours; each neighbour's distance byte is read ONCE and classifies
// it for good — unreached or level+1: successor (claim it, see below, if
// unreached); level-1: predecessor — because when a level-l vertex is expanded all distances
// <= l are final and every unreached neighbour is about to become l+1.  The classification
// (successor mask, predecessor count) is stored per queue position, so the backward pass needs
// no distance reads at all for slab neighbours and skips the row fetch of BFS leaves.
// Backward (thread per vertex): bk = 1 + sum of q over the successors in CSR order (the first two
// come from the per-position successor slots, so most vertices need no row at all), q = bk / npred; the
// vertex's distance byte is overwritten with a 'finished' mark that carries npred.  Then a coalesced
// sweep adds q[v] * npred[v] to the CTA's partial load.
// Claims are PLAIN STORES (no shared-memory atomics): every thread that reads a neighbour as unreached
// stores the new distance and appends it.  Two threads racing for the same vertex inside the few
// cycles between read and store both append it; such a duplicate queue entry is harmless — expanding
// it finds nothing new, its classification and its q value are identical, and the load is added per
// VERTEX by the sweep — and it is rare (a few per cent), whereas the atomics cost ~2 LSU cycles per lane
// on every claim.  The queue has room for 2n entries; a source that overflows it is handed back.
// Vertices of degree > kMidHubDeg are expanded by a whole warp (forward: deferred list;
// backward: lanes over adjacency entries, sum still in CSR order); rows longer than the slab
// continue in the CSR arrays with distance-based classification.
// Sources whose BFS is deeper than 126 levels, that meet a vertex with > 62 predecessors or whose queue
// (2n entries) overflows
// are reported back (status) and re-run by the batched engine; nothing is approximated.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nec {

constexpr int kMidThreads = 512;          // two CTAs per SM: one source's narrow levels overlap another's wide ones
constexpr uint32_t kMidMaxN = 180000;     // 1.125 B/vertex of shared memory + staging + bookkeeping <= 227 KB
constexpr uint32_t kMidMinN = 64;         // one source per CTA wins down to tiny graphs (3.5x at n = 300..2048, scripts/exp_small_n.py)
constexpr uint32_t kMidHubDeg = 64;
constexpr int kMidHubCap = 512;
constexpr uint32_t kMidMaxLevel = 126;    // distance bytes < 0x80; the backward pass overwrites them with a 'finished' mark
constexpr uint32_t kMidMaxPred = 62;      // finished mark = 0x80 | (level parity << 6) | npred  (0xFF stays 'unreached')
constexpr bool kMidAtomicClaims = false;   // false: plain-store claims, harmless duplicate queue entries (see below)
constexpr uint32_t kClsRowFlag = 1u << 15;   // cls: bits 0..14 successor mask, bit 15 "row needed anyway", bits 16..31 npred

enum MidStatus : uint8_t { kMidPending = 0, kMidDone = 1, kMidTooDeep = 2, kMidPredOverflow = 3, kMidQueueOverflow = 4 };

struct MidParams {
    uint32_t n, n_pad;                        // n_pad: n rounded up to 16
    const uint32_t *__restrict__ row_ptr;
    const uint32_t *__restrict__ col_idx;
    const uint4 *__restrict__ slab;           // n rows of W u32: [degree, first W-1 neighbours] (see nec_graph)
    const uint8_t *__restrict__ deg8;         // min(degree, 255) per vertex (255: look at row_ptr)
    const uint32_t *__restrict__ sources;     // nullptr: source index si is vertex source_base + si
    uint32_t source_base;
    uint32_t n_sources;
    int include_endpoints;
    double *q_base;        size_t q_stride;
    uint32_t *order_base;  size_t order_stride;   // stride == capacity of the per-CTA queue (>= n)
    uint4 *pos4_base;                         // per queue position {cls, first three BFS successors}; same stride as order
    double *bslot_base;    size_t bslot_stride;
    // measure_only != 0: forward pass only; per source index eccentricity / sum of distances / reached
    // count (nec_distance_measures) — they fall out of the level offsets, no extra traversal work
    int measure_only;
    uint32_t *out_ecc;
    unsigned long long *out_sumdist;
    uint32_t *out_reached;
    uint8_t *source_status;                   // per source index
    unsigned long long *counters;             // [0] reached [1] edge pairs [2] visits [3] max depth [4..6] cycles
};

// dynamic shared memory: dist | (claim bitmap, only with kMidAtomicClaims) | slab staging (chunk-major: [W/4][threads] x 16 B)
constexpr int kMidThreadsSmall = 128;     // small graphs: eight 128-thread CTAs per SM (levels hold a few hundred vertices at most;
                                          // what counts is how many sources are in flight, not threads per source)
__host__ __device__ inline size_t mid_smem_bytes(uint32_t n, int slab_w, int threads = kMidThreads) {
    const size_t n_pad = ((size_t)n + 15) / 16 * 16;
    const size_t claim = kMidAtomicClaims ? (((size_t)n + 31) / 32 * 4 + 15) / 16 * 16 : 0;
    return n_pad + claim + (size_t)threads * slab_w * 4;
}

// Adjacency "slab": one aligned row of W 32-bit words per vertex = its degree followed by its first
// W-1 neighbours in adjacency order.  One 32 B (W=8) or 64 B (W=16) transfer replaces the dependent
// row_ptr -> col_idx pair of loads for all low-degree vertices.
// Staging: thread t's 16-byte chunk k lives at stage[k * kMidThreads + t], so a warp's cp.async
// writes and LDS.128 reads are 512 contiguous bytes (conflict-free).
template <int W, int kT>
__device__ __forceinline__ void slab_prefetch(uint4 *stage_t, const uint4 *__restrict__ slab, uint32_t v) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(stage_t);
    const uint4 *src = slab + (size_t)v * (W / 4);
#pragma unroll
    for (int i = 0; i < W / 4; ++i)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * kT * i), "l"(src + i) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Predicated shared-memory accesses on 32-bit shared addresses (no branches, no generic addressing).
__device__ __forceinline__ uint32_t lds_u8_if(bool pred, uint32_t saddr, uint32_t otherwise) {
    uint32_t v = otherwise;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u8 %0, [%1];\n\t}" : "+r"(v) : "r"(saddr), "r"((uint32_t)pred) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8_if(bool pred, uint32_t saddr, uint32_t value) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t}" ::"r"(saddr), "r"(value), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ uint32_t atoms_or_if(bool pred, uint32_t saddr, uint32_t value) {
    uint32_t old = 0xFFFFFFFFu;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.or.b32 %0, [%1], %2;\n\t}" : "+r"(old) : "r"(saddr), "r"(value), "r"((uint32_t)pred) : "memory");
    return old;
}

template <int W, int kT = kMidThreads>
__global__ void __launch_bounds__(kT, kT == kMidThreads ? 2 : 8)
mid_engine(const MidParams p) {
    extern __shared__ __align__(16) unsigned char mid_smem[];
    constexpr int kWarps = kT / 32;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    constexpr uint32_t kStride = kT;
    constexpr uint32_t kSlabNbrs = W - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t n = p.n;
    uint8_t *dist = mid_smem;
    uint32_t *claim = reinterpret_cast<uint32_t *>(dist + p.n_pad);
    const uint32_t claim_words = kMidAtomicClaims ? (n + 31) / 32 : 0;
    const uint32_t dist_s32 = (uint32_t)__cvta_generic_to_shared(dist);
    const uint32_t claim_s32 = (uint32_t)__cvta_generic_to_shared(claim);
    uint4 *stage = reinterpret_cast<uint4 *>(mid_smem + p.n_pad + (((size_t)claim_words * 4 + 15) / 16 * 16)) + threadIdx.x;

    __shared__ uint32_t s_lvl[kMidMaxLevel + 3];
    __shared__ uint32_t s_tail, s_flag, s_hub_n, s_qfull;
    __shared__ unsigned long long s_msum;
    __shared__ uint32_t s_mreach;
    const uint32_t qcap = (uint32_t)p.order_stride;
    __shared__ uint32_t s_hub[kMidHubCap];
    __shared__ unsigned long long s_cnt[3];      // reached, edge pairs, queue entries (incl. duplicates)

    double *__restrict__ q = p.q_base + (size_t)blockIdx.x * p.q_stride;
    uint32_t *__restrict__ order = p.order_base + (size_t)blockIdx.x * p.order_stride;
    uint4 *__restrict__ pos4 = p.pos4_base + (size_t)blockIdx.x * p.order_stride;
    double *__restrict__ bslot = p.bslot_base + (size_t)blockIdx.x * p.bslot_stride;

    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    if (!p.measure_only) {                        // this CTA's partial load starts at zero (ordered before the first sweep by the barriers of the first source)
        double2 *b2 = reinterpret_cast<double2 *>(bslot);
        for (uint32_t i = threadIdx.x; i < (n + 1) / 2; i += kT) __stcs(&b2[i], make_double2(0.0, 0.0));
    }
    uint32_t my_depth = 0;
    long long cyc_reset = 0, cyc_fwd = 0, cyc_bwd = 0, cyc_small = 0;

    auto read_stage = [&](uint32_t (&w)[W], bool valid) {
#pragma unroll
        for (int k = 0; k < W / 4; ++k) {
            const uint4 t = valid ? stage[(size_t)k * kT] : make_uint4(0u, 0u, 0u, 0u);
            w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
        }
    };
    // always true (ids and degrees are < n), but makes the re-arming of a staging slot depend on
    // every shared-memory read of the previous contents having landed in registers
    auto landed = [&](const uint32_t (&w)[W]) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < W / 4; ++k) ok = ok && w[4 * k] != kNone;
        return ok;
    };

    for (uint32_t si = blockIdx.x; si < p.n_sources; si += gridDim.x) {
        const uint32_t s = p.sources ? p.sources[si] : p.source_base + si;
        long long t_phase = clock64();
        uint32_t t_reached = 0, t_edges = 0;      // per thread, this source
        // ---- reset (shared memory only) ------------------------------------------------
        {
            uint4 *d16 = reinterpret_cast<uint4 *>(dist);
            const uint4 inf4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            for (uint32_t i = threadIdx.x; i < p.n_pad / 16; i += kT) d16[i] = inf4;
            for (uint32_t i = threadIdx.x; i < claim_words; i += kT) claim[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            dist[s] = 0;
            if (kMidAtomicClaims) claim[s >> 5] = 1u << (s & 31);
            order[0] = s;
            s_tail = 1; s_flag = 0; s_hub_n = 0; s_qfull = 0; s_msum = 0; s_mreach = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_reset += t - t_phase; t_phase = t; }

        // ---- forward ---------------------------------------------------------------------
        auto claim_vertex = [&](uint32_t x, uint32_t next_d) -> bool {     // x is known to read as unreached
            bool won = true;
            if (kMidAtomicClaims) {
                const uint32_t bit = 1u << (x & 31);
                won = !(atomicOr(&claim[x >> 5], bit) & bit);
            }
            if (won) dist[x] = (uint8_t)next_d;
            return won;
        };
        auto try_claim = [&](uint32_t x, uint32_t next_d) -> bool {        // x == kNone: empty slot
            const bool cand = x != kNone && dist[x] == 0xFFu;
            return cand && claim_vertex(x, next_d);
        };
        auto append = [&](bool found, uint32_t x) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, found);
            if (bal) {
                uint32_t base = 0;
                const int leader = __ffs(bal) - 1;
                if (lane == leader) base = atomicAdd(&s_tail, (uint32_t)__popc(bal));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (base + __popc(bal) > qcap) { if (lane == leader) s_qfull = 1; }
                else if (found) __stcs(&order[base + __popc(bal & lane_lt)], x);
            }
        };

        uint32_t level = 0, lb = 0, le = 0;
        bool too_deep = false;
        for (;;) {
            __syncthreads();
            const uint32_t tail_now = min(s_tail, qcap);
            const uint32_t hubs = min(s_hub_n, (uint32_t)kMidHubCap);
            __syncthreads();
            // hubs deferred by the level that just ran belong to `level - 1`: expand them now,
            // before the frontier of `level` is frozen
            if (hubs) {
                const uint32_t next_d = level;            // they were on level-1
                for (uint32_t h = warp; h < hubs; h += kWarps) {
                    const uint32_t u = s_hub[h];
                    const uint32_t rs = __ldg(p.row_ptr + u), re = __ldg(p.row_ptr + u + 1);
                    for (uint32_t e0 = rs; e0 < re; e0 += 32) {
                        const uint32_t e = e0 + lane;
                        const uint32_t x = e < re ? __ldg(p.col_idx + e) : kNone;
                        append(try_claim(x, next_d), x);
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) s_hub_n = 0;
                continue;                                 // re-sample the tail
            }
            if (s_qfull) break;                           // queue overflowed: some slots below the tail were never
                                                          // written, so nothing behind this point may be expanded
            lb = le;
            le = tail_now;
            if (threadIdx.x == 0) s_lvl[level] = lb;
            if (le == lb) break;
            if (level >= kMidMaxLevel) { too_deep = true; break; }
            const uint32_t next_d = level + 1, pred_d = level - 1;   // level 0: pred_d wraps, matches nothing
            const long long t_lvl = clock64();
            const bool small_lvl = le - lb <= kStride;

            // software pipeline over this thread's frontier vertices: the vertex id two steps ahead
            // is being loaded, the slab row one step ahead is in flight to shared memory (cp.async)
            uint32_t i = lb + threadIdx.x;
            uint32_t v_cur = i < le ? __ldcs(&order[i]) : kNone;
            uint32_t v_nxt = i + kStride < le ? __ldcs(&order[i + kStride]) : kNone;
            if (v_cur != kNone) slab_prefetch<W, kT>(stage, p.slab, v_cur);
            cp_async_commit();
            for (uint32_t wbase = lb + warp * 32; wbase < le; wbase += kStride, i += kStride) {
                uint32_t row[W];
                cp_async_wait_all();
                read_stage(row, v_cur != kNone);
                const uint32_t u = v_cur;
                const uint32_t v_nn = i + 2 * kStride < le ? __ldcs(&order[i + 2 * kStride]) : kNone;
                if (v_nxt != kNone && landed(row)) slab_prefetch<W, kT>(stage, p.slab, v_nxt);
                cp_async_commit();
                v_cur = v_nxt; v_nxt = v_nn;

                uint32_t deg = 0, cls = 0;
                if (u != kNone) {
                    deg = row[0];
                    if (deg > kSlabNbrs) cls = kClsRowFlag;
                    if (deg > kMidHubDeg) {
                        const uint32_t pos = atomicAdd(&s_hub_n, 1u);
                        if (pos < (uint32_t)kMidHubCap) { s_hub[pos] = u; deg = 0; }   // else: expand inline (slow, correct)
                    }
                }
                // read each slab neighbour's distance once: classify it, claim it if unreached.
                // Branch-free (predicated shared-memory PTX): lanes differ per slot, branches would serialise.
                uint32_t won = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0, nsucc = 0;
#pragma unroll
                for (int k = 1; k < W; ++k) {
                    const bool valid = (uint32_t)(k - 1) < deg;
                    const uint32_t x = row[k];
                    const uint32_t d = lds_u8_if(valid, dist_s32 + x, 0xFFFFu);
                    const bool unreached = d == 0xFFu;
                    const bool is_succ = unreached || d == next_d;
                    s3 = (is_succ && nsucc == 3) ? x : s3;
                    s2 = (is_succ && nsucc == 2) ? x : s2;
                    s1 = (is_succ && nsucc == 1) ? x : s1;
                    s0 = (is_succ && nsucc == 0) ? x : s0;
                    nsucc += is_succ ? 1u : 0u;
                    bool mine = unreached;
                    if (kMidAtomicClaims) {
                        const uint32_t bit = 1u << (x & 31);
                        const uint32_t old = atoms_or_if(unreached, claim_s32 + ((x >> 5) << 2), bit);   // ~0 if not executed
                        mine = !(old & bit);
                    }
                    sts_u8_if(mine, dist_s32 + x, next_d);
                    won |= (mine ? 1u : 0u) << k;
                    cls |= (is_succ ? 1u : 0u) << (k - 1);
                    cls += (d == pred_d ? 1u : 0u) << 16;
                }
                if (u != kNone)                       // {s0 | npred<<26, s1 | min(nsucc,7)<<28 | rowflag<<31, s2, s3}; ids < 2^18
                    __stcs(&pos4[i], make_uint4(s0 | ((cls >> 16) << 26), s1 | (min(nsucc, 7u) << 28) | ((cls & kClsRowFlag) << 16), s2, s3));
                {   // append everything this warp found with ONE tail update
                    const uint32_t cnt = __popc(won);
                    const uint32_t incl = warp_inclusive_scan(cnt, lane);
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    if (total) {
                        uint32_t base = 0;
                        if (lane == 31) base = atomicAdd(&s_tail, total);
                        base = __shfl_sync(0xFFFFFFFFu, base, 31);
                        if (base + total > qcap) {
                            if (lane == 31) s_qfull = 1;
                        } else {
                            base += incl - cnt;
#pragma unroll
                            for (int k = 1; k < W; ++k)
                                if ((won >> k) & 1u) __stcs(&order[base++], row[k]);
                        }
                    }
                }
                // rows longer than the slab continue in the CSR arrays
                const uint32_t maxdeg = __reduce_max_sync(0xFFFFFFFFu, deg);
                if (maxdeg > kSlabNbrs) {
                    const uint32_t rs = deg > kSlabNbrs ? __ldg(p.row_ptr + u) : 0u;
                    for (uint32_t k0 = kSlabNbrs; k0 < maxdeg; k0 += 4) {
                        uint32_t xs[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) xs[j] = (k0 + j < deg) ? __ldg(p.col_idx + rs + k0 + j) : kNone;
#pragma unroll
                        for (int j = 0; j < 4; ++j) append(try_claim(xs[j], next_d), xs[j]);
                    }
                }
            }
            cp_async_wait_all();
            if (threadIdx.x == 0 && small_lvl) cyc_small += clock64() - t_lvl;
            ++level;
        }
        __syncthreads();
        if (too_deep || s_qfull) {
            if (threadIdx.x == 0) p.source_status[si] = too_deep ? kMidTooDeep : kMidQueueOverflow;
            __syncthreads();
            continue;
        }
        const uint32_t depth = level - 1;
        if (threadIdx.x == 0) { const long long t = clock64(); cyc_fwd += t - t_phase; t_phase = t; }
        if (p.measure_only) {
            // exact per-vertex sums from the distance vector (queue entries may contain duplicates)
            unsigned long long sum = 0;
            uint32_t cnt = 0;
            for (uint32_t v = threadIdx.x; v < n; v += kT) {
                const uint32_t d = dist[v];
                if (d != 0xFFu) {
                    sum += d; cnt += 1;
                    const uint32_t dg = __ldg(p.deg8 + v);
                    t_edges += dg != 0xFFu ? dg : __ldg(p.row_ptr + v + 1) - __ldg(p.row_ptr + v);
                }
            }
            t_reached = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o); cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); }
            if (lane == 0) { atomicAdd(&s_msum, sum); atomicAdd(&s_mreach, cnt); }
            {
                const uint32_t wr = __reduce_add_sync(0xFFFFFFFFu, t_reached), we = __reduce_add_sync(0xFFFFFFFFu, t_edges);
                if (lane == 0) { atomicAdd(&s_cnt[0], (unsigned long long)wr); atomicAdd(&s_cnt[1], (unsigned long long)we); }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (p.out_ecc) p.out_ecc[si] = depth;
                if (p.out_sumdist) p.out_sumdist[si] = s_msum;
                if (p.out_reached) p.out_reached[si] = s_mreach;
                p.source_status[si] = kMidDone;
            }
            if (depth > my_depth) my_depth = depth;
            __syncthreads();
            continue;
        }

        // ---- backward: deepest level first; one thread per vertex, CSR-order sums -----------
        // Everything a vertex needs was recorded per queue position by the forward pass: its first four
        // successors, successor count and npred, 16 B per position.  The streams are read coalesced, one step
        // ahead; only the q pulls are random.  Vertices with > 4 successors, rows longer than the slab
        // and hubs (rare) read the slab row / CSR and classify by distance byte: a finished vertex
        // carries 0x80 | (its level's parity << 6) | npred; a neighbour of a level-l vertex is a
        // successor <=> it is finished with the parity of l+1 (finished same-level neighbours carry l's).
        for (uint32_t l = depth; l >= 1; --l) {
            const uint32_t b0 = s_lvl[l], b1 = s_lvl[l + 1];
            const uint32_t d_pred = l - 1;
            const uint32_t succ_tag = 0x80u | (((l + 1) & 1u) << 6), mark_tag = 0x80u | ((l & 1u) << 6);
            auto is_successor_byte = [&](uint32_t b) { return (b & 0xC0u) == succ_tag && b != 0xFFu; };
            const long long t_lvl = clock64();
            // two-deep software pipeline: step t+2's (vertex, record) are being loaded, step t+1's q pulls
            // are in flight, step t is consumed
            uint32_t i = b0 + threadIdx.x;
            uint32_t v_c = kNone, m_c = 0, v_n = kNone;        // m: np<<8 | nsucc<<1 | rowflag of the step
            uint4 p_n = make_uint4(0u, 0u, 0u, 0u);
            double q0_c = 0.0, q1_c = 0.0, q2_c = 0.0, q3_c = 0.0;
            auto issue_pulls = [&](const uint4 &r, double &a0, double &a1, double &a2, double &a3) -> uint32_t {
                const uint32_t ns = (r.y >> 28) & 7u;
                a0 = ns >= 1 ? q[r.x & 0x3FFFFFFu] : 0.0;
                a1 = ns >= 2 ? q[r.y & 0xFFFFFFFu] : 0.0;
                a2 = ns >= 3 ? q[r.z] : 0.0;
                a3 = ns >= 4 ? q[r.w] : 0.0;
                return ((r.x >> 26) << 8) | (ns << 1) | (r.y >> 31);
            };
            if (i < b1) {
                v_c = __ldcs(&order[i]);
                const uint4 t = __ldcs(&pos4[i]);
                m_c = issue_pulls(t, q0_c, q1_c, q2_c, q3_c);
            }
            if (i + kStride < b1) { v_n = __ldcs(&order[i + kStride]); p_n = __ldcs(&pos4[i + kStride]); }
            for (uint32_t wbase = b0 + warp * 32; wbase < b1; wbase += kStride, i += kStride) {
                const uint32_t w = v_c, meta = m_c;
                double acc = ((q0_c + q1_c) + q2_c) + q3_c;                      // successors in CSR order
                // issue step t+1's pulls, then step t+2's loads
                v_c = v_n;
                m_c = issue_pulls(p_n, q0_c, q1_c, q2_c, q3_c);
                v_n = kNone; p_n = make_uint4(0u, 0u, 0u, 0u);
                if (i + 2 * kStride < b1) { v_n = __ldcs(&order[i + 2 * kStride]); p_n = __ldcs(&pos4[i + 2 * kStride]); }
                const uint32_t nsucc = (meta >> 1) & 7u;
                uint32_t np = meta >> 8;
                uint32_t deg = 0, rs = 0;
                bool hub = false;
                if (nsucc > 4 || (meta & 1u)) {                               // rare: the row itself is needed
                    uint32_t row[W];
#pragma unroll
                    for (int k = 0; k < W / 4; ++k) {
                        const uint4 t = __ldg(p.slab + (size_t)w * (W / 4) + k);
                        row[4 * k] = t.x; row[4 * k + 1] = t.y; row[4 * k + 2] = t.z; row[4 * k + 3] = t.w;
                    }
                    deg = row[0];
                    hub = deg > kMidHubDeg;
                    if (!hub) {
                        // recompute the whole sum in CSR order, classifying by distance byte
                        acc = 0.0;
#pragma unroll
                        for (int k = 1; k < W; ++k) {
                            const uint32_t b = (uint32_t)(k - 1) < deg ? (uint32_t)dist[row[k]] : 0xFFu;
                            if (is_successor_byte(b)) acc += q[row[k]];
                        }
                        if (deg > kSlabNbrs) {
                            rs = __ldg(p.row_ptr + w);
                            for (uint32_t k0 = kSlabNbrs; k0 < deg; k0 += 4) {
                                uint32_t xs[4], dy[4];
                                double qy[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) xs[j] = (k0 + j < deg) ? __ldg(p.col_idx + rs + k0 + j) : kNone;
#pragma unroll
                                for (int j = 0; j < 4; ++j) dy[j] = xs[j] != kNone ? (uint32_t)dist[xs[j]] : 0xFFu;
#pragma unroll
                                for (int j = 0; j < 4; ++j) qy[j] = is_successor_byte(dy[j]) ? q[xs[j]] : 0.0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) { acc += qy[j]; np += dy[j] == d_pred ? 1u : 0u; }
                            }
                        }
                    } else {
                        rs = __ldg(p.row_ptr + w);
                    }
                }
                // hubs of this warp's 32 vertices: all lanes cooperate, one hub at a time; the sum over
                // each 32-entry chunk is taken lane by lane so the order is still CSR order.  Hubs ignore
                // cls (the forward pass defers them before classifying).
                uint32_t hub_mask = __ballot_sync(0xFFFFFFFFu, hub);
                while (hub_mask) {
                    const int owner = __ffs(hub_mask) - 1;
                    hub_mask &= hub_mask - 1;
                    const uint32_t h_rs = __shfl_sync(0xFFFFFFFFu, rs, owner);
                    const uint32_t h_deg = __shfl_sync(0xFFFFFFFFu, deg, owner);
                    double h_acc = 0.0;
                    uint32_t h_np = 0;
                    for (uint32_t e0 = 0; e0 < h_deg; e0 += 32) {
                        double v = 0.0;
                        bool is_pred = false;
                        if (e0 + lane < h_deg) {
                            const uint32_t x = __ldg(p.col_idx + h_rs + e0 + lane);
                            const uint32_t dxx = dist[x];
                            if (is_successor_byte(dxx)) v = q[x];
                            is_pred = dxx == d_pred;
                        }
                        h_np += __popc(__ballot_sync(0xFFFFFFFFu, is_pred));
                        const uint32_t live = __ballot_sync(0xFFFFFFFFu, v != 0.0);
                        for (uint32_t bits = live; bits; bits &= bits - 1)      // ascending lane = CSR order
                            h_acc += __shfl_sync(0xFFFFFFFFu, v, __ffs(bits) - 1);
                    }
                    if (lane == owner) { acc = h_acc; np = h_np; }
                }
                if (w != kNone) {
                    const double bk = 1.0 + acc;
                    q[w] = bk / (double)np;
                    if (np > kMidMaxPred) s_flag = 1;
                    dist[w] = (uint8_t)(mark_tag | (np & 0x3Fu));           // finished: the sweep reads npred from here
                }
            }
            if (threadIdx.x == 0 && b1 - b0 <= kStride) cyc_small += clock64() - t_lvl;
            __syncthreads();
        }
        const bool overflow = s_flag != 0;
        if (overflow) {
            if (threadIdx.x == 0) p.source_status[si] = kMidPredOverflow;
            __syncthreads();
            continue;
        }
        // ---- sweep: this source's contribution, coalesced ------------------------------------
        // two vertices per 16-byte access, four accesses in flight per thread: q (L2) and the partial
        // load (HBM stream) are the long-latency loads here
        {
            const double sub = p.include_endpoints ? 0.0 : 1.0;
            auto one = [&](uint32_t v, uint32_t d, uint32_t dg, double qv, double &b) {      // exact stats + load term
                if (d != 0xFFu) {
                    t_reached += 1;                                   // exact (the queue may hold duplicates)
                    t_edges += dg != 0xFFu ? dg : __ldg(p.row_ptr + v + 1) - __ldg(p.row_ptr + v);
                    // the finished mark carries npred, and q * npred is bk to within one unit in the last place
                    // (q = bk / npred was rounded once).  bk >= 1 always, and only bk == 1 (a BFS leaf) can come
                    // back below 1 (e.g. (1/49)*49): the clamp makes a leaf contribute exactly bk - 1 = 0, as
                    // `b += bk; b -= 1.0` does in the crate (generic_graph.rs:1371-1374).  No FMA contraction.
                    if (d != 0u) b += __dsub_rn(fmax(__dmul_rn(qv, (double)(d & 0x3Fu)), 1.0), sub);
                }
            };
            const uint32_t n2 = n & ~1u;
            for (uint32_t base = 2 * threadIdx.x; base < n2; base += 8 * kT) {
                double2 qq[4], bb[4];
                uint32_t dd[4], gg[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t v = base + u * 2 * kT;
                    dd[u] = v < n2 ? (uint32_t)*reinterpret_cast<const uint16_t *>(dist + v) : 0xFFFFu;
                    const bool any = dd[u] != 0xFFFFu;
                    qq[u] = any ? *reinterpret_cast<const double2 *>(q + v) : make_double2(0.0, 0.0);
                    bb[u] = any ? __ldcs(reinterpret_cast<const double2 *>(bslot + v)) : make_double2(0.0, 0.0);
                    gg[u] = any ? (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(p.deg8 + v)) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t v = base + u * 2 * kT;
                    if (dd[u] != 0xFFFFu) {
                        one(v, dd[u] & 0xFFu, gg[u] & 0xFFu, qq[u].x, bb[u].x);
                        one(v + 1, dd[u] >> 8, gg[u] >> 8, qq[u].y, bb[u].y);
                        __stcs(reinterpret_cast<double2 *>(bslot + v), bb[u]);
                    }
                }
            }
            if ((n & 1u) && threadIdx.x == 0) {
                const uint32_t v = n - 1;
                double b = bslot[v];
                one(v, dist[v], p.deg8[v], q[v], b);
                bslot[v] = b;
            }
        }
        if (depth > my_depth) my_depth = depth;
        {
            const uint32_t wr = __reduce_add_sync(0xFFFFFFFFu, t_reached), we = __reduce_add_sync(0xFFFFFFFFu, t_edges);
            if (lane == 0) { atomicAdd(&s_cnt[0], (unsigned long long)wr); atomicAdd(&s_cnt[1], (unsigned long long)we); }
        }
        if (threadIdx.x == 0) { p.source_status[si] = kMidDone; cyc_bwd += clock64() - t_phase; s_cnt[2] += s_lvl[depth + 1]; }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&p.counters[0], s_cnt[0]);
        atomicAdd(&p.counters[1], s_cnt[1]);
        atomicAdd(&p.counters[2], s_cnt[2] ? s_cnt[2] : s_cnt[0]);
        atomicMax(&p.counters[3], (unsigned long long)my_depth);
        atomicAdd(&p.counters[4], (unsigned long long)cyc_reset);
        atomicAdd(&p.counters[5], (unsigned long long)cyc_fwd);
        atomicAdd(&p.counters[6], (unsigned long long)cyc_bwd);
        atomicAdd(&p.counters[7], (unsigned long long)cyc_small);
    }
}

}  // namespace nec
