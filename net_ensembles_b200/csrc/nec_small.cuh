// One-CTA-per-graph vertex_load for batches of small graphs (nec_vertex_load_batch): the
// shape of a Markov chain that measures after every step (reference:
// benches/common/mod.rs:79-90 — m_steps_quiet(k) then vertex_load(true), N = 50..~500).
//
// All mutable state of a graph lives in shared memory: dist[n][32] (u16), q[n][32] (f64),
// two f64 vectors for the per-batch and the running load.  The CTA walks the graph's
// ceil(n/32) source batches one after the other; lane k of every warp owns source k of the
// batch (same source-minor layout and the same per-(vertex,lane) arithmetic, lane-sum tree
// and accumulation order as the large-graph engine in nec_kernels.cuh, so results are
// bit-identical to nec_vertex_load for these sizes).  Levels are topology-driven (every
// warp tests its vertices' rows for the current level; n is tiny), no queues, no atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/net_ensembles_cuda.h"
#include "nec_kernels.cuh"

namespace nec {

constexpr int kSmallThreads = 256;
constexpr uint32_t kSmallInf = 0xFFFFu;

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
};

__host__ __device__ inline size_t small_smem_bytes(uint32_t n) {
    return (size_t)n * (kLanes * sizeof(double) + 2 * sizeof(double) + kLanes * sizeof(uint16_t));
}

__global__ void __launch_bounds__(kSmallThreads)
small_graph_kernel(const SmallParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kWarps = kSmallThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_more;

    unsigned long long my_reached = 0, my_edges = 0, my_visits = 0;
    uint32_t my_depth = 0;

    for (uint32_t g = blockIdx.x; g < p.n_graphs; g += gridDim.x) {
        const uint32_t n = p.n_per_graph[g];
        const uint32_t *__restrict__ rp = p.rp + p.rp_off[g];
        const uint32_t *__restrict__ col = p.col + p.col_off[g];
        double *q = reinterpret_cast<double *>(smem_raw);
        double *bsum = q + (size_t)n * kLanes;
        double *bpart = bsum + n;
        uint16_t *dist = reinterpret_cast<uint16_t *>(bpart + n);

        for (uint32_t v = threadIdx.x; v < n; v += kSmallThreads) bsum[v] = 0.0;

        const uint32_t n_batches = (n + kLanes - 1) / kLanes;
        for (uint32_t b = 0; b < n_batches; ++b) {
            __syncthreads();
            {   // reset (n*32 u16 = n*16 u32 words)
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dist);
                for (uint32_t i = threadIdx.x; i < n * (kLanes / 2); i += kSmallThreads) d32[i] = 0xFFFFFFFFu;
                for (uint32_t v = threadIdx.x; v < n; v += kSmallThreads) bpart[v] = 0.0;
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t s = b * kLanes + lane;
                if (s < n) dist[(size_t)s * kLanes + lane] = 0;
            }
            // ---- forward: expand level by level until nothing new is found ---------------
            uint32_t level = 0;
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) s_more = 0;
                __syncthreads();
                bool found = false;
                for (uint32_t u = warp; u < n; u += kWarps) {
                    const bool active = dist[(size_t)u * kLanes + lane] == level;
                    const uint32_t am = __ballot_sync(0xFFFFFFFFu, active);
                    if (am == 0) continue;
                    const uint32_t rs = __ldg(rp + u), re = __ldg(rp + u + 1);
                    if (lane == 0) {
                        my_reached += __popc(am);
                        my_edges += (unsigned long long)__popc(am) * (re - rs);
                        my_visits += 1;
                    }
                    for (uint32_t e = rs; e < re; ++e) {
                        const uint32_t x = __ldg(col + e);
                        if (active && dist[(size_t)x * kLanes + lane] == kSmallInf) {
                            dist[(size_t)x * kLanes + lane] = (uint16_t)(level + 1);
                            found = true;
                        }
                    }
                }
                if (__any_sync(0xFFFFFFFFu, found) && lane == 0) s_more = 1;
                __syncthreads();
                if (!s_more) break;
                ++level;
            }
            const uint32_t depth = level;
            if (depth > my_depth) my_depth = depth;
            // ---- backward: deepest level first, pull from successors ---------------------
            for (uint32_t l = depth; l >= 1; --l) {
                for (uint32_t w = warp; w < n; w += kWarps) {
                    const bool active = dist[(size_t)w * kLanes + lane] == l;
                    if (__ballot_sync(0xFFFFFFFFu, active) == 0) continue;
                    const uint32_t rs = __ldg(rp + w), re = __ldg(rp + w + 1);
                    double acc = 0.0;
                    uint32_t npred = 0;
                    for (uint32_t e = rs; e < re; ++e) {       // CSR order: fixes the rounding
                        const uint32_t x = __ldg(col + e);
                        const uint32_t dx = dist[(size_t)x * kLanes + lane];
                        if (active && dx == l + 1) acc += q[(size_t)x * kLanes + lane];
                        npred += (dx == l - 1) ? 1u : 0u;
                    }
                    const double bk = 1.0 + acc;
                    double contrib = 0.0;
                    if (active) {
                        q[(size_t)w * kLanes + lane] = bk / (double)npred;
                        contrib = p.include_endpoints ? bk : bk - 1.0;
                    }
                    const double total = warp_sum(contrib);
                    if (lane == 0) bpart[w] += total;
                }
                __syncthreads();
            }
            __syncthreads();
            for (uint32_t v = threadIdx.x; v < n; v += kSmallThreads) bsum[v] += bpart[v];
        }
        __syncthreads();
        double *out = p.out + p.out_off[g];
        for (uint32_t v = threadIdx.x; v < n; v += kSmallThreads) out[v] = bsum[v];
        __syncthreads();
    }
    if (lane == 0) {
        atomicAdd(&p.counters[0], my_reached);
        atomicAdd(&p.counters[1], my_edges);
        atomicAdd(&p.counters[2], my_visits);
        atomicMax(&p.counters[3], (unsigned long long)my_depth);
    }
}

// Device staging for nec_vertex_load_batch, owned by the context and grown on demand.
struct SmallWorkspace {
    void *base = nullptr;
    size_t cap = 0;
    bool attr_set = false;
    void release() { if (base) cudaFree(base); base = nullptr; cap = 0; }
};

inline int small_fail(std::string *err, int code, const std::string &msg) { if (err) *err = msg; return code; }

inline int small_batch_run(SmallWorkspace &ws, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, std::string *err,
                           nec_stats *stats, uint32_t n_graphs, const uint32_t *n_per_graph, const uint64_t *rp_off,
                           const uint32_t *rp, const uint64_t *col_off, const uint32_t *col, int include_endpoints, double *out) {
    if (n_graphs == 0) return NEC_OK;
    // ---- validate + size ------------------------------------------------------------------
    uint64_t rp_total = 0, col_total = 0, out_total = 0;
    uint32_t max_n = 0;
    std::vector<uint64_t> out_off(n_graphs);
    for (uint32_t g = 0; g < n_graphs; ++g) {
        const uint32_t n = n_per_graph[g];
        if (n == 0 || n > NEC_BATCH_MAX_N)
            return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: graph " + std::to_string(g) + " has n=" + std::to_string(n) +
                                                   " (allowed 1.." + std::to_string(NEC_BATCH_MAX_N) + "; use nec_vertex_load for larger graphs)");
        const uint32_t *r = rp + rp_off[g];
        if (r[0] != 0) return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: local row_ptr must start at 0 (graph " + std::to_string(g) + ")");
        for (uint32_t v = 0; v < n; ++v)
            if (r[v + 1] < r[v]) return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: row_ptr not monotone (graph " + std::to_string(g) + ")");
        const uint32_t *c = col + col_off[g];
        for (uint32_t e = 0; e < r[n]; ++e)
            if (c[e] >= n) return small_fail(err, NEC_EINVAL, "nec_vertex_load_batch: neighbour id out of range (graph " + std::to_string(g) + ")");
        rp_total = std::max<uint64_t>(rp_total, rp_off[g] + n + 1);
        col_total = std::max<uint64_t>(col_total, col_off[g] + r[n]);
        out_off[g] = out_total;
        out_total += n;
        max_n = std::max(max_n, n);
        stats->n_sources += n;
        stats->n_batches += (n + kLanes - 1) / kLanes;
    }
    // ---- device staging: one arena -----------------------------------------------------------
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_n = up((size_t)n_graphs * 4), b_off = up((size_t)n_graphs * 8);
    const size_t b_rp = up(rp_total * 4), b_col = up(std::max<uint64_t>(col_total, 1) * 4), b_out = up(out_total * 8);
    const size_t need = b_n + 3 * b_off + b_rp + b_col + b_out + 256;
    if (need > ws.cap) {
        ws.release();
        cudaError_t e = cudaMalloc(&ws.base, need);
        if (e != cudaSuccess) return small_fail(err, NEC_ENOMEM, std::string("nec_vertex_load_batch: cudaMalloc failed: ") + cudaGetErrorString(e));
        ws.cap = need;
    }
    uint8_t *d = (uint8_t *)ws.base;
    uint32_t *d_n = (uint32_t *)d; d += b_n;
    uint64_t *d_rp_off = (uint64_t *)d; d += b_off;
    uint64_t *d_col_off = (uint64_t *)d; d += b_off;
    uint64_t *d_out_off = (uint64_t *)d; d += b_off;
    uint32_t *d_rp = (uint32_t *)d; d += b_rp;
    uint32_t *d_col = (uint32_t *)d; d += b_col;
    double *d_out = (double *)d; d += b_out;
    unsigned long long *d_cnt = (unsigned long long *)d;
    cudaMemcpyAsync(d_n, n_per_graph, (size_t)n_graphs * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_rp_off, rp_off, (size_t)n_graphs * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_col_off, col_off, (size_t)n_graphs * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_out_off, out_off.data(), (size_t)n_graphs * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_rp, rp, rp_total * 4, cudaMemcpyHostToDevice, stream);
    if (col_total) cudaMemcpyAsync(d_col, col, col_total * 4, cudaMemcpyHostToDevice, stream);
    cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), stream);

    const size_t smem = small_smem_bytes(max_n);
    cudaError_t e = cudaFuncSetAttribute(small_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(NEC_BATCH_MAX_N));
    if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    SmallParams p{};
    p.n_graphs = n_graphs; p.n_per_graph = d_n; p.rp_off = d_rp_off; p.rp = d_rp; p.col_off = d_col_off; p.col = d_col;
    p.out_off = d_out_off; p.include_endpoints = include_endpoints; p.out = d_out; p.counters = d_cnt;
    cudaEventRecord(ev0, stream);
    small_graph_kernel<<<n_graphs, kSmallThreads, smem, stream>>>(p);
    cudaEventRecord(ev1, stream);
    e = cudaGetLastError();
    if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: launch failed: ") + cudaGetErrorString(e));
    unsigned long long cnt[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(out, d_out, out_total * 8, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, stream);
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return small_fail(err, NEC_ECUDA, std::string("nec_vertex_load_batch: ") + cudaGetErrorString(e));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    stats->kernel_ms = ms;
    stats->engine_ms = ms;
    stats->kernel_launches = 1;
    stats->engine = 3;
    stats->reached_pairs = cnt[0]; stats->edge_pairs = cnt[1]; stats->vertex_visits = cnt[2]; stats->max_depth = (uint32_t)cnt[3];
    stats->n_slots = n_graphs;
    stats->workspace_bytes = ws.cap;
    return NEC_OK;
}

}  // namespace nec
