// Device-resident Markov chains: the adjacency lists of many small graphs stay in HBM
// between measurements, in slack-padded rows, and every Markov step arrives as the handful of LIST OPERATIONS
// the host performed on its copy (push / swap_remove / retarget — the exact primitives of the crate's
// containers: graph.rs:101-142, sw_graph.rs:259-279, :350-442), replayed in order by one warp per graph.
// Adjacency ORDER on the device therefore stays identical to the host's (it defines the rounding of the load),
// and a measurement step ships a few bytes per chain instead of re-exporting and re-uploading every CSR
// (reference shape: benches/common/mod.rs:79-90 — m_steps_quiet(k), then vertex_load(true)).
// The load itself is the one-CTA-per-graph kernel of nec_small.cuh reading the padded rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/net_ensembles_cuda.h"
#include "nec_small.cuh"

namespace nec {

enum ChainFlag : uint32_t { kChainRowOverflow = 1u, kChainBadOp = 2u };

// one CTA per graph: CSR (local ids) -> padded rows of `cap` u16 entries, degree per vertex
__global__ void chain_fill_kernel(uint32_t n_graphs, const uint32_t *__restrict__ n_per, const uint64_t *__restrict__ rp_off,
                                  const uint32_t *__restrict__ rp, const uint64_t *__restrict__ col_off, const uint32_t *__restrict__ col,
                                  const uint64_t *__restrict__ vert_base, uint32_t cap, uint16_t *__restrict__ rows, uint16_t *__restrict__ deg) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    for (uint32_t g = blockIdx.x; g < n_graphs; g += gridDim.x) {
        const uint32_t n = n_per[g];
        const uint32_t *rp_g = rp + rp_off[g];
        const uint32_t *col_g = col + col_off[g];
        uint16_t *rows_g = rows + vert_base[g] * cap;
        uint16_t *deg_g = deg + vert_base[g];
        for (uint32_t v = warp; v < n; v += warps) {
            const uint32_t rs = rp_g[v], d = rp_g[v + 1] - rs;
            for (uint32_t k = lane; k < cap; k += 32) rows_g[(size_t)v * cap + k] = k < d ? (uint16_t)col_g[rs + k] : (uint16_t)0xFFFFu;
            if (lane == 0) deg_g[v] = (uint16_t)d;
        }
    }
}

// one warp per graph: replay the graph's list operations in order.  op = kind << 60 | v << 40 | x << 20 | y
__global__ void chain_apply_kernel(uint32_t n_graphs, const uint32_t *__restrict__ n_per, const uint64_t *__restrict__ vert_base, uint32_t cap,
                                   uint16_t *__restrict__ rows, uint16_t *__restrict__ deg, const uint32_t *__restrict__ op_off,
                                   const unsigned long long *__restrict__ ops, uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n_graphs) return;
    const uint32_t n = n_per[g];
    uint16_t *rows_g = rows + vert_base[g] * cap;
    uint16_t *deg_g = deg + vert_base[g];
    uint32_t bad = 0;
    for (uint32_t i = op_off[g]; i < op_off[g + 1]; ++i) {
        const unsigned long long op = ops[i];
        const uint32_t kind = (uint32_t)(op >> 60), v = (uint32_t)(op >> 40) & 0xFFFFFu, x = (uint32_t)(op >> 20) & 0xFFFFFu, y = (uint32_t)op & 0xFFFFFu;
        if (v >= n || x >= n || y >= n || kind < 1 || kind > 3) { bad |= kChainBadOp; continue; }
        uint16_t *row = rows_g + (size_t)v * cap;
        const uint32_t d = deg_g[v];
        if (kind == 1) {                                            // push x
            if (d >= cap) { bad |= kChainRowOverflow; continue; }
            if (lane == 0) { row[d] = (uint16_t)x; deg_g[v] = (uint16_t)(d + 1); }
        } else {
            // position of x in v's list (entries are unique): lanes scan the row
            uint32_t pos = 0xFFFFFFFFu;
            for (uint32_t k0 = 0; k0 < d && pos == 0xFFFFFFFFu; k0 += 32) {
                const uint32_t k = k0 + lane;
                const uint32_t hit = __ballot_sync(0xFFFFFFFFu, k < d && row[k] == (uint16_t)x);
                if (hit) pos = k0 + (uint32_t)__ffs(hit) - 1;
            }
            if (pos == 0xFFFFFFFFu) { bad |= kChainBadOp; continue; }
            if (lane == 0) {
                if (kind == 2) { row[pos] = row[d - 1]; row[d - 1] = (uint16_t)0xFFFFu; deg_g[v] = (uint16_t)(d - 1); }   // swap_remove
                else row[pos] = (uint16_t)y;                                                                      // retarget
            }
        }
        __syncwarp();
    }
    if (bad && lane == 0) atomicOr(&flags[g], bad);
}

}  // namespace nec

// The opaque handle of the C ABI (nec_chains_*).
struct nec_chains {
    uint32_t n_graphs = 0, cap = 0, max_n = 0;
    uint64_t n_vertices = 0;                      // sum of the graphs' vertex counts = length of the result
    std::vector<uint32_t> h_n;
    std::vector<uint32_t> h_nnz;                  // adjacency entries per graph, kept in step with the replayed operations
    std::vector<uint64_t> h_vert_base;            // n_graphs + 1
    void *d_base = nullptr; size_t d_bytes = 0;
    uint32_t *d_n = nullptr, *d_flags = nullptr, *d_op_off = nullptr;
    uint64_t *d_vert_base = nullptr;
    uint16_t *d_rows = nullptr, *d_deg = nullptr;
    double *d_out = nullptr;
    unsigned long long *d_ops = nullptr; size_t d_ops_cap = 0;
    unsigned long long *d_cnt = nullptr;
    // pinned host side: op staging, result, flags + counters
    void *h_base = nullptr; size_t h_bytes = 0;
    double *h_out = nullptr;
    uint32_t *h_flags = nullptr, *h_op_off = nullptr;
    unsigned long long *h_cnt = nullptr;
    unsigned long long *h_ops = nullptr; size_t h_ops_cap = 0;
    uint64_t steps_applied = 0;
};

namespace nec {

inline void chains_release(nec_chains *c) {
    if (!c) return;
    if (c->d_base) cudaFree(c->d_base);
    if (c->d_ops) cudaFree(c->d_ops);
    if (c->h_base) cudaFreeHost(c->h_base);
    if (c->h_ops) cudaFreeHost(c->h_ops);
    delete c;
}

inline int chains_fail(std::string *err, int code, const std::string &msg) { if (err) *err = msg; return code; }

// Build the device-resident set from a batch of CSRs (same input format as nec_vertex_load_batch).
inline int chains_create(SmallWorkspace &ws, cudaStream_t stream, std::string *err, uint32_t n_graphs, const uint32_t *n_per_graph,
                         const uint64_t *rp_off, const uint32_t *rp, const uint64_t *col_off, const uint32_t *col, uint32_t slack,
                         nec_chains **out) {
    *out = nullptr;
    uint32_t max_deg = 0, max_n = 0;
    uint64_t n_vertices = 0, rp_total = 0, col_total = 0;
    for (uint32_t g = 0; g < n_graphs; ++g) {
        const uint32_t n = n_per_graph[g];
        if (n == 0 || n > NEC_BATCH_MAX_N)
            return chains_fail(err, NEC_EINVAL, "nec_chains_create: graph " + std::to_string(g) + " has n=" + std::to_string(n) + " (allowed 1.." + std::to_string(NEC_BATCH_MAX_N) + ")");
        const uint32_t *r = rp + rp_off[g];
        if (r[0] != 0) return chains_fail(err, NEC_EINVAL, "nec_chains_create: local row_ptr must start at 0 (graph " + std::to_string(g) + ")");
        for (uint32_t v = 0; v < n; ++v) {
            if (r[v + 1] < r[v]) return chains_fail(err, NEC_EINVAL, "nec_chains_create: row_ptr not monotone (graph " + std::to_string(g) + ")");
            max_deg = std::max(max_deg, r[v + 1] - r[v]);
        }
        const uint32_t *cc = col + col_off[g];
        for (uint32_t k = 0; k < r[n]; ++k)
            if (cc[k] >= n) return chains_fail(err, NEC_EINVAL, "nec_chains_create: neighbour id out of range (graph " + std::to_string(g) + ")");
        max_n = std::max(max_n, n);
        n_vertices += n;
        rp_total = std::max<uint64_t>(rp_total, rp_off[g] + n + 1);
        col_total = std::max<uint64_t>(col_total, col_off[g] + r[n]);
    }
    nec_chains *c = new (std::nothrow) nec_chains();
    if (!c) return chains_fail(err, NEC_ENOMEM, "out of host memory");
    c->n_graphs = n_graphs; c->max_n = max_n; c->n_vertices = n_vertices;
    // row capacity: room for the lists to grow (a chain's degrees fluctuate around a fixed mean); a row that
    // outgrows it is reported and the caller re-creates the set with more slack
    c->cap = std::min<uint32_t>(NEC_BATCH_MAX_N, ((std::max<uint32_t>(max_deg, 4u) + std::max<uint32_t>(slack, 4u) + 7u) / 8u) * 8u);
    c->h_n.assign(n_per_graph, n_per_graph + n_graphs);
    c->h_nnz.resize(n_graphs);
    for (uint32_t g = 0; g < n_graphs; ++g) c->h_nnz[g] = (rp + rp_off[g])[n_per_graph[g]];
    c->h_vert_base.resize((size_t)n_graphs + 1);
    c->h_vert_base[0] = 0;
    for (uint32_t g = 0; g < n_graphs; ++g) c->h_vert_base[g + 1] = c->h_vert_base[g] + n_per_graph[g];
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t b_n = up((size_t)n_graphs * 4), b_vb = up(((size_t)n_graphs + 1) * 8), b_rows = up(n_vertices * c->cap * 2), b_deg = up(n_vertices * 2),
                 b_out = up(n_vertices * 8), b_flags = up((size_t)n_graphs * 4), b_opoff = up(((size_t)n_graphs + 1) * 4), b_cnt = 256;
    c->d_bytes = b_n + b_vb + b_rows + b_deg + b_out + b_flags + b_opoff + b_cnt;
    cudaError_t e = cudaMalloc(&c->d_base, c->d_bytes);
    if (e != cudaSuccess) { chains_release(c); return chains_fail(err, NEC_ENOMEM, std::string("nec_chains_create: cudaMalloc failed: ") + cudaGetErrorString(e)); }
    {
        uint8_t *d = (uint8_t *)c->d_base;
        c->d_n = (uint32_t *)d; d += b_n;
        c->d_vert_base = (uint64_t *)d; d += b_vb;
        c->d_rows = (uint16_t *)d; d += b_rows;
        c->d_deg = (uint16_t *)d; d += b_deg;
        c->d_out = (double *)d; d += b_out;
        c->d_flags = (uint32_t *)d; d += b_flags;
        c->d_op_off = (uint32_t *)d; d += b_opoff;
        c->d_cnt = (unsigned long long *)d;
    }
    c->h_bytes = b_out + b_flags + b_opoff + b_cnt;
    e = cudaHostAlloc(&c->h_base, c->h_bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { chains_release(c); return chains_fail(err, NEC_ENOMEM, std::string("nec_chains_create: cudaHostAlloc failed: ") + cudaGetErrorString(e)); }
    {
        uint8_t *h = (uint8_t *)c->h_base;
        c->h_out = (double *)h; h += b_out;
        c->h_flags = (uint32_t *)h; h += b_flags;
        c->h_op_off = (uint32_t *)h; h += b_opoff;
        c->h_cnt = (unsigned long long *)h;
    }
    // one-off upload of the CSR batch through a temporary device buffer, then the fill kernel
    void *tmp = nullptr;
    const size_t t_rpoff = up((size_t)n_graphs * 8), t_rp = up(rp_total * 4), t_col = up(std::max<uint64_t>(col_total, 1) * 4);
    e = cudaMalloc(&tmp, 2 * t_rpoff + t_rp + t_col);
    if (e != cudaSuccess) { chains_release(c); return chains_fail(err, NEC_ENOMEM, std::string("nec_chains_create: cudaMalloc failed: ") + cudaGetErrorString(e)); }
    uint8_t *t = (uint8_t *)tmp;
    uint64_t *t_rpo = (uint64_t *)t; t += t_rpoff;
    uint64_t *t_colo = (uint64_t *)t; t += t_rpoff;
    uint32_t *t_rpv = (uint32_t *)t; t += t_rp;
    uint32_t *t_colv = (uint32_t *)t;
    e = cudaMemcpyAsync(c->d_n, n_per_graph, (size_t)n_graphs * 4, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_vert_base, c->h_vert_base.data(), ((size_t)n_graphs + 1) * 8, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t_rpo, rp_off, (size_t)n_graphs * 8, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t_colo, col_off, (size_t)n_graphs * 8, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t_rpv, rp, rp_total * 4, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && col_total) e = cudaMemcpyAsync(t_colv, col, col_total * 4, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_flags, 0, (size_t)n_graphs * 4, stream);
    if (e == cudaSuccess) {
        chain_fill_kernel<<<std::min<uint32_t>(n_graphs, 148u * 8u), 256, 0, stream>>>(n_graphs, c->d_n, t_rpo, t_rpv, t_colo, t_colv, c->d_vert_base, c->cap, c->d_rows, c->d_deg);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { chains_release(c); return chains_fail(err, NEC_ECUDA, std::string("nec_chains_create: ") + cudaGetErrorString(e)); }
    (void)ws;
    *out = c;
    return NEC_OK;
}

// Apply the step's list operations (op_offsets[g] .. op_offsets[g+1] belong to graph g; may be all empty), then
// vertex_load of every graph.  out == NULL leaves the result in the set's pinned buffer (nec_chains_result).
inline int chains_step_and_load(SmallWorkspace &ws, cudaStream_t stream, std::string *err, nec_stats *stats, nec_chains *c,
                                const uint32_t *op_offsets, const uint64_t *ops, int include_endpoints, double *out) {
    const uint32_t n_graphs = c->n_graphs;
    if (n_graphs == 0) return NEC_OK;
    cudaError_t e = ws.init_pipeline();
    if (e != cudaSuccess) return chains_fail(err, NEC_ECUDA, std::string("nec_chains: stream/event setup failed: ") + cudaGetErrorString(e));
    const uint32_t n_ops = op_offsets ? op_offsets[n_graphs] : 0;
    cudaEventRecord(ws.ev_entry, stream);
    cudaStreamWaitEvent(ws.s_in, ws.ev_entry, 0);
    cudaStreamWaitEvent(ws.s_k, ws.ev_entry, 0);
    cudaStreamWaitEvent(ws.s_k2, ws.ev_entry, 0);
    if (n_ops) {
        if (!ops) return chains_fail(err, NEC_EINVAL, "nec_chains_step: ops is NULL");
        if (n_ops > c->h_ops_cap) {
            if (c->h_ops) cudaFreeHost(c->h_ops);
            if (c->d_ops) cudaFree(c->d_ops);
            c->h_ops = nullptr; c->d_ops = nullptr; c->h_ops_cap = c->d_ops_cap = 0;
            const size_t cap = (size_t)n_ops * 2 + 1024;
            if (cudaHostAlloc((void **)&c->h_ops, cap * 8, cudaHostAllocDefault) != cudaSuccess || cudaMalloc((void **)&c->d_ops, cap * 8) != cudaSuccess)
                return chains_fail(err, NEC_ENOMEM, "nec_chains_step: cannot allocate the operation staging");
            c->h_ops_cap = c->d_ops_cap = cap;
        }
        std::memcpy(c->h_ops, ops, (size_t)n_ops * 8);
        std::memcpy(c->h_op_off, op_offsets, ((size_t)n_graphs + 1) * 4);
        for (uint32_t g = 0; g < n_graphs; ++g)          // list lengths after this step (push +1, swap_remove -1)
            for (uint32_t i = op_offsets[g]; i < op_offsets[g + 1]; ++i) {
                const uint64_t kind = ops[i] >> 60;
                if (kind == 1) c->h_nnz[g] += 1;
                else if (kind == 2 && c->h_nnz[g]) c->h_nnz[g] -= 1;
            }
        cudaMemcpyAsync(c->d_ops, c->h_ops, (size_t)n_ops * 8, cudaMemcpyHostToDevice, ws.s_in);
        cudaMemcpyAsync(c->d_op_off, c->h_op_off, ((size_t)n_graphs + 1) * 4, cudaMemcpyHostToDevice, ws.s_in);
        chain_apply_kernel<<<(n_graphs * 32 + 255) / 256, 256, 0, ws.s_in>>>(n_graphs, c->d_n, c->d_vert_base, c->cap, c->d_rows, c->d_deg, c->d_op_off, c->d_ops, c->d_flags);
        c->steps_applied++;
    }
    cudaMemsetAsync(c->d_cnt, 0, 4 * sizeof(unsigned long long), ws.s_in);
    cudaMemcpyAsync(c->h_flags, c->d_flags, (size_t)n_graphs * 4, cudaMemcpyDeviceToHost, ws.s_in);
    cudaEventRecord(ws.ev_in[0], ws.s_in);

    // shared memory reserved for a graph's compact CSR (row pointers u32 + neighbours u16): what the largest graph
    // needs now, capped; a graph beyond the cap is read from its padded rows through L1/L2 instead
    size_t max_csr = 0;
    for (uint32_t g = 0; g < n_graphs; ++g) max_csr = std::max(max_csr, (size_t)(c->h_n[g] + 1) * 4 + (size_t)c->h_nnz[g] * 2);
    const uint32_t csr_budget = (uint32_t)std::min<size_t>((max_csr + 255) / 256 * 256, kSmallCsrBytes);
    const size_t smem = small_smem_bytes(c->max_n, csr_budget);
    e = cudaFuncSetAttribute(small_graph_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(NEC_BATCH_MAX_N, kSmallCsrBytes));
    if (e != cudaSuccess) return chains_fail(err, NEC_ECUDA, std::string("nec_chains: cudaFuncSetAttribute: ") + cudaGetErrorString(e));

    const uint32_t n_chunks = std::max<uint32_t>(1, std::min<uint32_t>(kSmallMaxChunks, n_graphs / 256));
    for (uint32_t k = 0; k < n_chunks; ++k) {
        const uint32_t g0 = (uint32_t)((uint64_t)n_graphs * k / n_chunks), g1 = (uint32_t)((uint64_t)n_graphs * (k + 1) / n_chunks);
        cudaStream_t sk = (k & 1) ? ws.s_k2 : ws.s_k;
        if (k < 2) cudaStreamWaitEvent(sk, ws.ev_in[0], 0);
        SmallParams p{};
        p.n_graphs = g1 - g0; p.n_per_graph = c->d_n + g0; p.out_off = c->d_vert_base + g0; p.include_endpoints = include_endpoints;
        p.out = c->d_out; p.counters = c->d_cnt; p.csr_budget = csr_budget;
        p.ch_rows = c->d_rows; p.ch_deg = c->d_deg; p.ch_cap = c->cap;
        cudaEventRecord(ws.ev_k0[k], sk);
        small_graph_kernel<true><<<g1 - g0, kSmallThreads, smem, sk>>>(p);
        cudaEventRecord(ws.ev_k1[k], sk);
        cudaStreamWaitEvent(ws.s_out, ws.ev_k1[k], 0);
        const uint64_t o0 = c->h_vert_base[g0], o1 = c->h_vert_base[g1];
        cudaMemcpyAsync(c->h_out + o0, c->d_out + o0, (o1 - o0) * 8, cudaMemcpyDeviceToHost, ws.s_out);
        cudaEventRecord(ws.ev_out[k], ws.s_out);
    }
    e = cudaGetLastError();
    // hand the results over chunk by chunk while later chunks are still in flight
    for (uint32_t k = 0; k < n_chunks && e == cudaSuccess; ++k) {
        const uint32_t g0 = (uint32_t)((uint64_t)n_graphs * k / n_chunks), g1 = (uint32_t)((uint64_t)n_graphs * (k + 1) / n_chunks);
        e = cudaEventSynchronize(ws.ev_out[k]);
        if (e != cudaSuccess || !out) continue;
        const uint64_t o0 = c->h_vert_base[g0], o1 = c->h_vert_base[g1];
        const int64_t n_blk = (int64_t)((o1 - o0 + 16383) / 16384);
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < n_blk; ++b) {
            const uint64_t lo = o0 + (uint64_t)b * 16384, hi = std::min<uint64_t>(o1, lo + 16384);
            std::memcpy(out + lo, c->h_out + lo, (hi - lo) * 8);
        }
    }
    if (e == cudaSuccess) {
        cudaMemcpyAsync(c->h_cnt, c->d_cnt, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ws.s_out);
        e = cudaStreamSynchronize(ws.s_out);
    }
    cudaStreamSynchronize(ws.s_in);
    cudaStreamSynchronize(ws.s_k);
    cudaStreamSynchronize(ws.s_k2);
    if (e != cudaSuccess) return chains_fail(err, NEC_ECUDA, std::string("nec_chains_step: ") + cudaGetErrorString(e));
    for (uint32_t g = 0; g < n_graphs; ++g)
        if (c->h_flags[g])
            return chains_fail(err, c->h_flags[g] & kChainBadOp ? NEC_EINVAL : NEC_EINTERNAL,
                               std::string("nec_chains_step: graph ") + std::to_string(g) +
                                   (c->h_flags[g] & kChainBadOp ? ": a list operation does not apply (vertex out of range or entry not in the list); the device copy is out of step with the host's"
                                                                : ": an adjacency list outgrew its row (capacity " + std::to_string(c->cap) + "); re-create the set with more slack"));
    float ms_span = 0.f;
    for (uint32_t k = 0; k < n_chunks; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ws.ev_k0[0], ws.ev_k1[k]);
        ms_span = std::max(ms_span, ms);
    }
    stats->n_sources = c->n_vertices;
    for (uint32_t g = 0; g < n_graphs; ++g) stats->n_batches += (c->h_n[g] + kLanes - 1) / kLanes;
    stats->kernel_ms = ms_span; stats->engine_ms = ms_span;
    stats->kernel_launches = n_chunks + (n_ops ? 1u : 0u);
    stats->engine = 3;
    stats->reached_pairs = c->h_cnt[0]; stats->edge_pairs = c->h_cnt[1]; stats->vertex_visits = c->h_cnt[2]; stats->max_depth = (uint32_t)c->h_cnt[3];
    stats->n_slots = n_graphs;
    stats->workspace_bytes = c->d_bytes;
    stats->cta_threads = kSmallThreads; stats->team_size = 1;
    return NEC_OK;
}

// Device adjacency back to the host as a CSR batch (compact, local row pointers back to back: graph g's n_g+1
// row pointers start at sum_{h<g}(n_h+1); neighbour ids follow in graph order) — the check that the replayed
// lists equal the host's, order included.
inline int chains_download(cudaStream_t stream, std::string *err, nec_chains *c, uint32_t *row_ptr_concat, uint32_t *col_idx_concat, uint64_t col_capacity, uint64_t *nnz_out) {
    std::vector<uint16_t> rows((size_t)c->n_vertices * c->cap), deg(c->n_vertices);
    cudaError_t e = cudaMemcpyAsync(rows.data(), c->d_rows, rows.size() * 2, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(deg.data(), c->d_deg, deg.size() * 2, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return chains_fail(err, NEC_ECUDA, std::string("nec_chains_download: ") + cudaGetErrorString(e));
    uint64_t at_rp = 0, at_col = 0;
    for (uint32_t g = 0; g < c->n_graphs; ++g) {
        uint32_t local = 0;
        for (uint32_t v = 0; v < c->h_n[g]; ++v) {
            row_ptr_concat[at_rp++] = local;
            const uint64_t gv = c->h_vert_base[g] + v;
            for (uint32_t k = 0; k < deg[gv]; ++k) {
                if (at_col >= col_capacity) return chains_fail(err, NEC_EINVAL, "nec_chains_download: col_idx buffer too small");
                col_idx_concat[at_col++] = rows[gv * c->cap + k];
            }
            local += deg[gv];
        }
        row_ptr_concat[at_rp++] = local;
    }
    if (nnz_out) *nnz_out = at_col;
    return NEC_OK;
}

}  // namespace nec
