// One-CTA-per-graph vertex_load for batches of small graphs (nec_vertex_load_batch): the
// shape of a Markov chain that measures after every step (reference:
// benches/common/mod.rs:79-90 — m_steps_quiet(k) then vertex_load(true), N = 50..~500).
//
// One CTA per graph, ONE WARP PER SOURCE: every warp owns a private slice of shared memory
// (dist u32[n], q f64[n], BFS order u16[n], partial load f64[n]) and runs whole sources —
// forward BFS (lane per frontier vertex, claim by shared-memory atomicCAS, ballot-aggregated
// queue append) and the backward pull (lane per vertex, successors summed in CSR order,
// q = bk / npred) — with __syncwarp only; no block barrier inside a source.  Warp w takes
// sources w, w+W, ... in order and the warps' partial loads are summed in warp order at the
// end, so the result is reproducible run to run.  Per-source values bk_s(v) are bit-identical
// to the batched engine's (same formula, same CSR order; the mid engine adds q * npred, which is
// bk to within one unit in the last place); the order in which sources are added differs as
// well, so the batch agrees with nec_vertex_load to rounding (<= 1e-12), not bitwise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/net_ensembles_cuda.h"
#include "nec_kernels.cuh"

namespace nec {

constexpr int kSmallThreads = 256;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr uint32_t kSmallInf = 0xFFFFFFFFu;

struct SmallParams {
    uint32_t n_graphs;
    const uint32_t *__restrict__ n_per_graph;
    const uint64_t *__restrict__ rp_off;
    const uint32_t *__restrict__ rp;
    const uint64_t *__restrict__ col_off;
    const uint32_t *__restrict__ col;
    const uint64_t *__restrict__ out_off;
    int include_endpoints;
    double *__restrict__ out;
    unsigned long long *counters;   // [0] reached pairs [1] edge pairs [2] visits [3] max depth
    uint32_t csr_budget;            // bytes of shared memory reserved for a graph's CSR in this launch
    // device-resident chains (nec_chain.cuh, kChain): padded adjacency rows instead of a CSR batch; graph i's
    // vertex v has its row at ch_rows[(out_off[i] + v) * ch_cap] and its degree at ch_deg[out_off[i] + v]
    const uint16_t *__restrict__ ch_rows;
    const uint16_t *__restrict__ ch_deg;
    uint32_t ch_cap;
};

// per warp: q f64[n] | bpart f64[n] | dist u32[n] | order u16[n]   (n rounded up to 4)
__host__ __device__ inline size_t small_warp_bytes(uint32_t n) {
    const size_t np = ((size_t)n + 3) / 4 * 4;
    return np * (8 + 8 + 4 + 2);
}
constexpr uint32_t kSmallCsrBytes = 24 * 1024;   // shared-memory budget for a graph's CSR (row_ptr u32 + neighbours u16)
__host__ __device__ inline size_t small_smem_bytes(uint32_t n, uint32_t csr_budget) { return small_warp_bytes(n) * kSmallWarps + csr_budget; }

// Predicated shared-memory compare-and-swap (one ATOMS.CAS under a predicate instead of a divergent branch around it:
// in the forward pass only ~5 of 32 lanes meet an unreached neighbour).  Returns 0 when not executed.
__device__ __forceinline__ uint32_t atoms_cas_if(bool pred, uint32_t saddr, uint32_t compare, uint32_t value) {
    uint32_t old = 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p atom.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
                 : "+r"(old) : "r"(saddr), "r"(compare), "r"(value), "r"((uint32_t)pred) : "memory");
    return old;
}

template <bool kChain>
__global__ void __launch_bounds__(kSmallThreads)
small_graph_kernel(const SmallParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane_lt = (1u << lane) - 1u;
    unsigned long long my_reached = 0, my_edges = 0;
    uint32_t my_depth = 0;

    for (uint32_t g = blockIdx.x; g < p.n_graphs; g += gridDim.x) {
        const uint32_t n = p.n_per_graph[g];
        const uint32_t np4 = (n + 3) / 4 * 4;
        // stage the graph's CSR in shared memory when it fits (it is read ~2 x n times per graph)
        uint32_t *s_rp = reinterpret_cast<uint32_t *>(smem_raw + (size_t)kSmallWarps * small_warp_bytes(n));
        uint16_t *s_col = reinterpret_cast<uint16_t *>(s_rp + n + 1);
        const uint32_t *__restrict__ rp_g = nullptr, *__restrict__ col_g = nullptr;
        const uint16_t *__restrict__ rows_g = nullptr, *__restrict__ deg_g = nullptr;
        const uint32_t cap = p.ch_cap;
        bool staged;
        if constexpr (kChain) {
            // padded rows -> compact CSR in shared memory: warp 0 scans the degrees, then every warp copies rows
            rows_g = p.ch_rows + p.out_off[g] * cap;
            deg_g = p.ch_deg + p.out_off[g];
            __shared__ uint32_t s_total;
            if (warp == 0) {
                uint32_t carry = 0;
                for (uint32_t v0 = 0; v0 < n; v0 += 32) {
                    const uint32_t v = v0 + lane;
                    const uint32_t d = v < n ? (uint32_t)deg_g[v] : 0u;
                    const uint32_t incl = warp_inclusive_scan(d, lane);
                    if (v < n && (size_t)(n + 1) * 4 <= p.csr_budget) s_rp[v] = carry + incl - d;
                    carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
                if (lane == 0) { s_total = carry; if ((size_t)(n + 1) * 4 <= p.csr_budget) s_rp[n] = carry; }
            }
            __syncthreads();
            staged = (size_t)(n + 1) * 4 + (size_t)s_total * 2 <= p.csr_budget;
            if (staged)
                for (uint32_t v = warp; v < n; v += kSmallWarps) {
                    const uint32_t rs = s_rp[v], d = s_rp[v + 1] - rs;
                    for (uint32_t k = lane; k < d; k += 32) s_col[rs + k] = rows_g[(size_t)v * cap + k];
                }
        } else {
            rp_g = p.rp + p.rp_off[g];
            col_g = p.col + p.col_off[g];
            const uint32_t nnz = __ldg(rp_g + n);
            staged = (size_t)(n + 1) * 4 + (size_t)nnz * 2 <= p.csr_budget;
            if (staged) {
                for (uint32_t v = threadIdx.x; v <= n; v += kSmallThreads) s_rp[v] = __ldg(rp_g + v);
                for (uint32_t e = threadIdx.x; e < nnz; e += kSmallThreads) s_col[e] = (uint16_t)__ldg(col_g + e);
            }
        }
        __syncthreads();
        // Everything below runs once per graph in one of two instantiations - CSR staged in shared memory (the normal
        // case) or read from global memory - so that no adjacency access of the hot loops carries a branch.
        auto run_graph = [&](auto staged_tag) {
        constexpr bool kStaged = decltype(staged_tag)::value;
        // rows are [row_begin(v), row_end(v)) of col_at(.)
        auto row_begin = [&](uint32_t v) -> uint32_t {
            if constexpr (kStaged) return s_rp[v];
            else if constexpr (kChain) return v * cap; else return __ldg(rp_g + v);
        };
        auto row_end = [&](uint32_t v) -> uint32_t {
            if constexpr (kStaged) return s_rp[v + 1];
            else if constexpr (kChain) return v * cap + (uint32_t)deg_g[v]; else return __ldg(rp_g + v + 1);
        };
        auto col_at = [&](uint32_t e) -> uint32_t {
            if constexpr (kStaged) return (uint32_t)s_col[e];
            else if constexpr (kChain) return (uint32_t)rows_g[e]; else return __ldg(col_g + e);
        };
        unsigned char *mine = smem_raw + (size_t)warp * small_warp_bytes(n);
        double *q = reinterpret_cast<double *>(mine);
        double *bpart = q + np4;
        uint32_t *dist = reinterpret_cast<uint32_t *>(bpart + np4);
        uint16_t *order = reinterpret_cast<uint16_t *>(dist + np4);
        const double sub = p.include_endpoints ? 0.0 : 1.0;
        const uint32_t dist_s = (uint32_t)__cvta_generic_to_shared(dist);

        for (uint32_t v = lane; v < n; v += 32) bpart[v] = 0.0;

        for (uint32_t s = warp; s < n; s += kSmallWarps) {
            for (uint32_t v = lane; v < n; v += 32) dist[v] = kSmallInf;
            __syncwarp();
            if (lane == 0) { dist[s] = 0; order[0] = (uint16_t)s; }
            __syncwarp();
            // ---- forward: level by level inside the warp -----------------------------------
            uint32_t lb = 0, le = 1, level = 0;
            uint32_t lvl_start = 0;                         // lane L keeps the queue position where level L starts (levels < 32)
            while (lb < le) {
                if ((uint32_t)lane == level) lvl_start = lb;
                uint32_t tail = le;
                for (uint32_t base = lb; base < le; base += 32) {
                    const uint32_t i = base + lane;
                    uint32_t rs = 0, deg = 0;
                    if (i < le) {
                        const uint32_t u = order[i];
                        rs = row_begin(u);
                        deg = row_end(u) - rs;
                        my_reached += 1;
                        my_edges += deg;
                    }
                    const uint32_t maxdeg = __reduce_max_sync(0xFFFFFFFFu, deg);
                    for (uint32_t k0 = 0; k0 < maxdeg; k0 += 2) {
                        uint32_t x[2];
                        bool cand[2], found[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            x[j] = k0 + j < deg ? col_at(rs + k0 + j) : 0u;
                            cand[j] = k0 + j < deg && dist[x[j]] == kSmallInf;
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            found[j] = atoms_cas_if(cand[j], dist_s + x[j] * 4u, kSmallInf, level + 1) == kSmallInf;
                            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, found[j]);
                            if (found[j]) order[tail + __popc(bal & lane_lt)] = (uint16_t)x[j];
                            tail += __popc(bal);
                        }
                    }
                }
                __syncwarp();
                lb = le; le = tail; ++level;
            }
            const uint32_t depth = level - 1;               // deepest non-empty level
            if (depth > my_depth) my_depth = depth;
            // ---- backward: walk the BFS order from the back, one level at a time -------------
            // order[] is sorted by level; level l occupies a contiguous range ending at `hi`
            uint32_t hi = le;                               // == number of reached vertices
            for (uint32_t l = depth; l >= 1; --l) {
                // the start of level l: entries [lo, hi) have dist == l.  Kept from the forward pass for the first 32
                // levels; deeper ones are found by stepping back through the order
                uint32_t lo = hi;
                if (l < 32u) lo = __shfl_sync(0xFFFFFFFFu, lvl_start, (int)l);
                else for (;;) {                             // step back in chunks of 32
                    const uint32_t probe = lo > 32 ? lo - 32 : 0;
                    const uint32_t j = probe + lane;
                    const bool same = j < lo && dist[order[j]] == l;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, same);
                    const uint32_t span = lo - probe;       // lanes [0, span) are valid
                    const uint32_t valid_mask = span >= 32 ? 0xFFFFFFFFu : ((1u << span) - 1u);
                    if (bal == valid_mask && probe > 0) { lo = probe; continue; }
                    // the level starts inside this chunk: first lane (from the top) that is not level l
                    const uint32_t not_l = ~bal & valid_mask;
                    lo = not_l ? probe + (32 - __clz(not_l)) : probe;
                    break;
                }
                for (uint32_t base = lo; base < hi; base += 32) {
                    const uint32_t i = base + lane;
                    if (i < hi) {
                        const uint32_t w = order[i];
                        const uint32_t rs = row_begin(w), re = row_end(w);
                        double acc = 0.0;
                        uint32_t npred = 0;
                        uint32_t e = rs;
                        for (; e + 2 <= re; e += 2) {                  // CSR order: fixes the rounding
                            const uint32_t x0 = col_at(e), x1 = col_at(e + 1);
                            const uint32_t d0 = dist[x0], d1 = dist[x1];
                            const double q0 = d0 == l + 1 ? q[x0] : 0.0, q1 = d1 == l + 1 ? q[x1] : 0.0;
                            acc += q0; acc += q1;
                            npred += (d0 == l - 1 ? 1u : 0u) + (d1 == l - 1 ? 1u : 0u);
                        }
                        if (e < re) {
                            const uint32_t x = col_at(e);
                            const uint32_t dx = dist[x];
                            if (dx == l + 1) acc += q[x];
                            npred += (dx == l - 1) ? 1u : 0u;
                        }
                        const double bk = 1.0 + acc;
                        q[w] = bk / (double)npred;
                        bpart[w] += bk - sub;
                    }
                }
                __syncwarp();
                hi = lo;
            }
            __syncwarp();
        }
        };                                              // run_graph
        if (staged) run_graph(std::true_type{});
        else run_graph(std::false_type{});
        __syncthreads();
        // ---- warps' partial loads, summed in warp order ------------------------------------------
        double *out = p.out + p.out_off[g];
        for (uint32_t v = threadIdx.x; v < n; v += kSmallThreads) {
            double acc = 0.0;
            for (int w = 0; w < kSmallWarps; ++w)
                acc += reinterpret_cast<const double *>(smem_raw + (size_t)w * small_warp_bytes(n))[np4 + v];
            out[v] = acc;
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_reached += __shfl_xor_sync(0xFFFFFFFFu, my_reached, o);
        my_edges += __shfl_xor_sync(0xFFFFFFFFu, my_edges, o);
    }
    if (lane == 0) {
        atomicAdd(&p.counters[0], my_reached);
        atomicAdd(&p.counters[1], my_edges);
        atomicAdd(&p.counters[2], my_reached);
        atomicMax(&p.counters[3], (unsigned long long)my_depth);
    }
}

// Staging for nec_vertex_load_batch, owned by the context and grown on demand: a device arena, a pinned
// host mirror of it, and the three streams of the copy-in / kernel / copy-out pipeline.
constexpr int kSmallMaxChunks = 8;
struct SmallWorkspace {
    void *base = nullptr;            // device arena
    size_t cap = 0;
    void *h_base = nullptr;          // pinned host staging (same layout)
    size_t h_cap = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_k2 = nullptr, s_out = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_in[kSmallMaxChunks] = {}, ev_k0[kSmallMaxChunks] = {}, ev_k1[kSmallMaxChunks] = {}, ev_out[kSmallMaxChunks] = {};
    bool attr_set = false;
    cudaError_t init_pipeline() {
        if (s_in) return cudaSuccess;
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&s_k2, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_entry, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int c = 0; c < kSmallMaxChunks; ++c) {
            if ((e = cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&ev_k0[c])) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&ev_k1[c])) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_out[c], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    void release() {
        if (base) cudaFree(base);
        if (h_base) cudaFreeHost(h_base);
        base = h_base = nullptr; cap = h_cap = 0;
        if (s_in) {
            cudaStreamDestroy(s_in); cudaStreamDestroy(s_k); cudaStreamDestroy(s_k2); cudaStreamDestroy(s_out); cudaEventDestroy(ev_entry);
            for (int c = 0; c < kSmallMaxChunks; ++c) { cudaEventDestroy(ev_in[c]); cudaEventDestroy(ev_k0[c]); cudaEventDestroy(ev_k1[c]); cudaEventDestroy(ev_out[c]); }
            s_in = s_k = s_k2 = s_out = nullptr;
        }
    }
};

inline int small_fail(std::string *err, int code, const std::string &msg) { if (err) *err = msg; return code; }

// The call is synchronous (host pointers in, host pointers out) but pipelined inside: the graphs are cut
// into up to 8 chunks; while the kernel works on chunk c, chunk c+1 is validated and packed into pinned
// memory by all host cores and copied in, and chunk c-1's loads are copied out and handed to the caller.
inline int small_batch_run(SmallWorkspace &ws, cudaStream_t stream, std::string *err,
                           nec_stats *stats, uint32_t n_graphs, const uint32_t *n_per_graph, const uint64_t *rp_off,
                           const uint32_t *rp, const uint64_t *col_off, const uint32_t *col, int include_endpoints, double *out) {
    if (n_graphs == 0) return NEC_OK;
    const bool verbose = getenv("NEC_SMALL_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_pack = 0.0, t_wait = 0.0, t_copy = 0.0;
    // ---- sizes; compact offsets of the packed copy ---------------------------------------------
    uint64_t rp_total = 0, col_total = 0, out_total = 0;
    uint32_t max_n = 0;
    size_t max_csr = 0;
    std::vector<uint64_t> c_rp_off(n_graphs + 1), c_col_off(n_graphs + 1), out_off(n_graphs + 1);
    for (uint32_t g = 0; g < n_graphs; ++g) {
        const uint32_t n = n_per_graph[g];
        if (n == 0 || n > NEC_BATCH_MAX_N)
            return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: graph " + std::to_string(g) + " has n=" + std::to_string(n) +
                                                   " (allowed 1.." + std::to_string(NEC_BATCH_MAX_N) + "; use nec_vertex_load for larger graphs)");
        const uint32_t nnz = (rp + rp_off[g])[n];
        if (nnz > (uint64_t)n * n)
            return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: row_ptr not monotone (graph " + std::to_string(g) + ")");
        c_rp_off[g] = rp_total; c_col_off[g] = col_total; out_off[g] = out_total;
        rp_total += n + 1; col_total += nnz; out_total += n;
        max_n = std::max(max_n, n);
        max_csr = std::max(max_csr, (size_t)(n + 1) * 4 + (size_t)nnz * 2);
        stats->n_sources += n;
        stats->n_batches += (n + kLanes - 1) / kLanes;
    }
    c_rp_off[n_graphs] = rp_total; c_col_off[n_graphs] = col_total; out_off[n_graphs] = out_total;

    // ---- staging: one device arena + its pinned mirror -----------------------------------------
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_n = up((size_t)n_graphs * 4), b_off = up((size_t)(n_graphs + 1) * 8);
    const size_t b_rp = up(rp_total * 4), b_col = up(std::max<uint64_t>(col_total, 1) * 4), b_out = up(out_total * 8);
    const size_t need = b_n + 3 * b_off + b_rp + b_col + b_out + 256;
    cudaError_t e = ws.init_pipeline();
    if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: stream/event setup failed: ") + cudaGetErrorString(e));
    if (need > ws.cap) {
        if (ws.base) cudaFree(ws.base);
        ws.base = nullptr; ws.cap = 0;
        e = cudaMalloc(&ws.base, need + need / 4);
        if (e != cudaSuccess) return small_fail(err, NEC_ENOMEM, std::string("nec_vertex_load_batch: cudaMalloc failed: ") + cudaGetErrorString(e));
        ws.cap = need + need / 4;
    }
    if (need > ws.h_cap) {
        if (ws.h_base) cudaFreeHost(ws.h_base);
        ws.h_base = nullptr; ws.h_cap = 0;
        e = cudaHostAlloc(&ws.h_base, need + need / 4, cudaHostAllocDefault);
        if (e != cudaSuccess) return small_fail(err, NEC_ENOMEM, std::string("nec_vertex_load_batch: cudaHostAlloc failed: ") + cudaGetErrorString(e));
        ws.h_cap = need + need / 4;
    }
    struct Planes { uint32_t *n; uint64_t *rp_off, *col_off, *out_off; uint32_t *rp, *col; double *out; unsigned long long *cnt; };
    auto carve = [&](void *base) {
        uint8_t *d = (uint8_t *)base;
        Planes pl{};
        pl.n = (uint32_t *)d; d += b_n;
        pl.rp_off = (uint64_t *)d; d += b_off;
        pl.col_off = (uint64_t *)d; d += b_off;
        pl.out_off = (uint64_t *)d; d += b_off;
        pl.rp = (uint32_t *)d; d += b_rp;
        pl.col = (uint32_t *)d; d += b_col;
        pl.out = (double *)d; d += b_out;
        pl.cnt = (unsigned long long *)d;
        return pl;
    };
    const Planes D = carve(ws.base), H = carve(ws.h_base);

    // work issued earlier on the context's stream stays ahead of this call
    cudaEventRecord(ws.ev_entry, stream);
    cudaStreamWaitEvent(ws.s_in, ws.ev_entry, 0);
    cudaStreamWaitEvent(ws.s_k, ws.ev_entry, 0);
    cudaStreamWaitEvent(ws.s_k2, ws.ev_entry, 0);
    std::memcpy(H.n, n_per_graph, (size_t)n_graphs * 4);
    std::memcpy(H.rp_off, c_rp_off.data(), (size_t)n_graphs * 8);
    std::memcpy(H.col_off, c_col_off.data(), (size_t)n_graphs * 8);
    std::memcpy(H.out_off, out_off.data(), (size_t)n_graphs * 8);
    cudaMemcpyAsync(D.n, H.n, b_n + 3 * b_off, cudaMemcpyHostToDevice, ws.s_in);      // the four planes are contiguous
    cudaMemsetAsync(D.cnt, 0, 4 * sizeof(unsigned long long), ws.s_in);    // ordered before every kernel via ev_in[]

    const uint32_t csr_budget = (uint32_t)std::min<size_t>((max_csr + 255) / 256 * 256, kSmallCsrBytes);   // larger CSRs are read from L1/L2
    const size_t smem = small_smem_bytes(max_n, csr_budget);
    {   // per device (a process may hold contexts on several): set on every call
        e = cudaFuncSetAttribute(small_graph_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(NEC_BATCH_MAX_N, kSmallCsrBytes));
        if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    }

    const uint32_t n_chunks = std::max<uint32_t>(1, std::min<uint32_t>(kSmallMaxChunks, n_graphs / 256));
    int bad_graph = -1, bad_kind = 0;
    uint32_t launched = 0;
    for (uint32_t c = 0; c < n_chunks && bad_graph < 0; ++c) {
        const uint32_t g0 = (uint32_t)((uint64_t)n_graphs * c / n_chunks), g1 = (uint32_t)((uint64_t)n_graphs * (c + 1) / n_chunks);
        // validate + pack this chunk's graphs on all host cores (graphs are independent)
        const double t_p0 = now();
#pragma omp parallel for schedule(static)
        for (int64_t g = g0; g < (int64_t)g1; ++g) {
            const uint32_t n = n_per_graph[g];
            const uint32_t *r = rp + rp_off[g];
            const uint32_t *cc = col + col_off[g];
            int kind = r[0] != 0 ? 1 : 0;
            for (uint32_t v = 0; v < n && !kind; ++v)
                if (r[v + 1] < r[v]) kind = 2;
            for (uint32_t k = 0; !kind && k < r[n]; ++k)
                if (cc[k] >= n) kind = 3;
            if (kind) {
#pragma omp critical
                { if (bad_graph < 0 || g < bad_graph) { bad_graph = (int)g; bad_kind = kind; } }
            } else {
                std::memcpy(H.rp + c_rp_off[g], r, (size_t)(n + 1) * 4);
                std::memcpy(H.col + c_col_off[g], cc, (size_t)r[n] * 4);
            }
        }
        t_pack += now() - t_p0;
        if (bad_graph >= 0) break;
        cudaMemcpyAsync(D.rp + c_rp_off[g0], H.rp + c_rp_off[g0], (c_rp_off[g1] - c_rp_off[g0]) * 4, cudaMemcpyHostToDevice, ws.s_in);
        if (c_col_off[g1] > c_col_off[g0])
            cudaMemcpyAsync(D.col + c_col_off[g0], H.col + c_col_off[g0], (c_col_off[g1] - c_col_off[g0]) * 4, cudaMemcpyHostToDevice, ws.s_in);
        cudaEventRecord(ws.ev_in[c], ws.s_in);
        // consecutive chunks' kernels go to alternating streams: the next chunk's CTAs fill the SMs that the
        // previous chunk's last wave leaves idle
        cudaStream_t sk = (c & 1) ? ws.s_k2 : ws.s_k;
        cudaStreamWaitEvent(sk, ws.ev_in[c], 0);
        SmallParams p{};
        p.n_graphs = g1 - g0; p.n_per_graph = D.n + g0; p.rp_off = D.rp_off + g0; p.rp = D.rp; p.col_off = D.col_off + g0; p.col = D.col;
        p.out_off = D.out_off + g0; p.include_endpoints = include_endpoints; p.out = D.out; p.counters = D.cnt;
        p.csr_budget = csr_budget;
        cudaEventRecord(ws.ev_k0[c], sk);
        small_graph_kernel<false><<<g1 - g0, kSmallThreads, smem, sk>>>(p);
        cudaEventRecord(ws.ev_k1[c], sk);
        cudaStreamWaitEvent(ws.s_out, ws.ev_k1[c], 0);
        cudaMemcpyAsync(H.out + out_off[g0], D.out + out_off[g0], (out_off[g1] - out_off[g0]) * 8, cudaMemcpyDeviceToHost, ws.s_out);
        cudaEventRecord(ws.ev_out[c], ws.s_out);
        launched = c + 1;
    }
    e = cudaGetLastError();
    const double t_issued = now();
    // ---- hand the results over chunk by chunk while later chunks are still in flight ----------
    for (uint32_t c = 0; c < launched && e == cudaSuccess; ++c) {
        const uint32_t g0 = (uint32_t)((uint64_t)n_graphs * c / n_chunks), g1 = (uint32_t)((uint64_t)n_graphs * (c + 1) / n_chunks);
        const double t_w0 = now();
        e = cudaEventSynchronize(ws.ev_out[c]);
        const double t_w1 = now();
        t_wait += t_w1 - t_w0;
        if (e != cudaSuccess || bad_graph >= 0) continue;
        const uint64_t o0 = out_off[g0], o1 = out_off[g1];
        const int64_t n_blk = (int64_t)((o1 - o0 + 16383) / 16384);
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < n_blk; ++b) {
            const uint64_t lo = o0 + (uint64_t)b * 16384, hi = std::min<uint64_t>(o1, lo + 16384);
            std::memcpy(out + lo, H.out + lo, (hi - lo) * 8);
        }
        t_copy += now() - t_w1;
    }
    if (verbose)
        fprintf(stderr, "[nec small] %u graphs %u chunks: setup+issue %.2f ms (pack %.2f) wait %.2f copy-out %.2f total %.2f\n",
                n_graphs, n_chunks, t_issued - t_begin, t_pack, t_wait, t_copy, now() - t_begin);
    unsigned long long cnt[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) {                       // every kernel has finished: its copy-out was waited for above
        cudaMemcpyAsync(H.cnt, D.cnt, sizeof cnt, cudaMemcpyDeviceToHost, ws.s_out);
        e = cudaStreamSynchronize(ws.s_out);
        std::memcpy(cnt, H.cnt, sizeof cnt);
    }
    cudaStreamSynchronize(ws.s_in);
    cudaStreamSynchronize(ws.s_k);
    cudaStreamSynchronize(ws.s_k2);
    cudaStreamSynchronize(ws.s_out);
    if (bad_graph >= 0)
        return small_fail(err, NEC_EINVAL, std::string("nec_vertex_load_batch: ") +
                                               (bad_kind == 1 ? "local row_ptr must start at 0" : bad_kind == 2 ? "row_ptr not monotone" : "neighbour id out of range") +
                                               " (graph " + std::to_string(bad_graph) + ")");
    if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: ") + cudaGetErrorString(e));
    float ms_span = 0.f;                          // first kernel's start to the last kernel's end (chunks overlap)
    for (uint32_t c = 0; c < launched; ++c) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ws.ev_k0[0], ws.ev_k1[c]);
        ms_span = std::max(ms_span, ms);
    }
    stats->kernel_ms = ms_span;
    stats->engine_ms = ms_span;
    stats->kernel_launches = launched;
    stats->engine = 3;
    stats->reached_pairs = cnt[0]; stats->edge_pairs = cnt[1]; stats->vertex_visits = cnt[2]; stats->max_depth = (uint32_t)cnt[3];
    stats->n_slots = n_graphs;
    stats->workspace_bytes = ws.cap;
    return NEC_OK;
}

}  // namespace nec
