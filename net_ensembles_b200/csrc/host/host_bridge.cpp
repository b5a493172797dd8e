// The drop-in call over a host-mirror handle (nec_handle.hpp), compiled into libnet_ensembles_cuda.so:
// <ensemble or graph>.vertex_load(include_endpoints) = export the adjacency lists to an order-preserving CSR,
// upload, run the CUDA path, return the load vector.  Reference call sites: src/traits/graph_traits.rs:328-330
// (any ensemble), benches/common/mod.rs:79-90 (Markov chain: m_steps_quiet then vertex_load(true)).
// No load arithmetic here: everything numeric happens behind the C ABI in nec_api.cu.
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "../../../include/net_ensembles_cuda.h"
#include "nec_handle.hpp"

using namespace nec_host;

extern "C" void nec_internal_set_error(nec_ctx *ctx, const char *msg);   // nec_api.cu (not part of the public header)

namespace {

template <class F>
int guarded(nec_ctx *ctx, F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        nec_internal_set_error(ctx, "out of host memory");
        return NEC_ENOMEM;
    } catch (const std::exception &e) {
        nec_internal_set_error(ctx, e.what());
        return NEC_EINVAL;
    }
}

struct ChainGroup {                          // one group's exported CSRs (buffers reused across calls)
    std::vector<uint32_t> n_per, rp, col;
    std::vector<uint64_t> rp_off, col_off;
    uint64_t out_count = 0;
    int bad = 0;
};
thread_local ChainGroup g_group;
thread_local nec_stats g_chain_stats;

void prepare_group(ChainGroup &grp, void *const *handles, uint32_t g0, uint32_t g1, uint64_t m_steps) {
    const uint32_t cnt = g1 - g0;
    grp.n_per.resize(cnt); grp.rp_off.resize(cnt); grp.col_off.resize(cnt);
    int bad = 0;
    std::vector<uint64_t> &nnz_per = grp.col_off;        // first the per-graph entry counts, then their prefix sums
    // chains are independent: step them on all host cores (each chain owns its generator).  A CSR's size is
    // 2 x edge_count (every edge sits in both endpoints' lists; the containers keep the count), so sizing needs
    // no walk over the 256 separately allocated adjacency vectors of each chain - the export below is the only
    // pass over them, and it checks the size it was promised
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int64_t g = g0; g < (int64_t)g1; ++g) {
        Handle *h = static_cast<Handle *>(handles[g]);
        for (uint64_t k = 0; k < m_steps; ++k)
            if (!h->m_step()) { bad += 1; break; }
        with_graph(h, [&](auto &gr) {
            grp.n_per[g - g0] = (uint32_t)gr.vertex_count();
            nnz_per[g - g0] = 2 * (uint64_t)gr.edge_count;
            return 0;
        });
    }
    grp.bad = bad;
    uint64_t rp_total = 0, col_total = 0, out_count = 0;
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint64_t nnz_k = nnz_per[k];
        grp.rp_off[k] = rp_total; grp.col_off[k] = col_total;
        rp_total += grp.n_per[k] + 1; col_total += nnz_k; out_count += grp.n_per[k];
    }
    grp.out_count = out_count;
    if (grp.rp.size() < rp_total) grp.rp.resize(rp_total + rp_total / 8);
    if (grp.col.size() < col_total + 1) grp.col.resize(col_total + col_total / 8 + 1);
    uint32_t *const rp_buf = grp.rp.data(), *const col_buf = grp.col.data();
    const uint64_t *const rp_off = grp.rp_off.data(), *const col_off = grp.col_off.data();
    int mismatch = 0;
#pragma omp parallel for schedule(static) reduction(+ : mismatch)
    for (int64_t g = g0; g < (int64_t)g1; ++g)
        with_graph(static_cast<Handle *>(handles[g]), [&](auto &gr) {
            uint32_t *r = rp_buf + rp_off[g - g0];
            uint32_t *c = col_buf + col_off[g - g0];
            const uint64_t room = 2 * (uint64_t)gr.edge_count;
            uint64_t at = 0;
            for (size_t v = 0; v < gr.vertices.size(); ++v) {
                r[v] = (uint32_t)at;
                const auto &adj = gr.vertices[v].adj;
                if (at + adj.size() > room) { mismatch += 1; break; }    // never: edge_count is maintained by add/remove
                for (auto &e : adj) c[at++] = e.to;
            }
            r[gr.vertices.size()] = (uint32_t)at;
            if (at != room) mismatch += 1;
            return 0;
        });
    if (mismatch) grp.bad = -2;
}

}  // namespace

namespace {
// adjacency lists of a handle -> device graph through the export fast path (the containers keep every vertex's
// neighbours in one contiguous slice; nec_graph_upload_adjacency flattens them into pinned staging)
template <class G>
int upload_graph(nec_ctx *ctx, const G &g, nec_graph **dg) {
    const size_t n = g.vertex_count();
    std::vector<const void *> rows(n);
    std::vector<uint64_t> deg(n);
    for (size_t v = 0; v < n; ++v) { rows[v] = g.vertices[v].adj.data(); deg[v] = g.vertices[v].adj.size(); }
    using Edge = typename std::decay<decltype(g.vertices[0].adj[0])>::type;
    return nec_graph_upload_adjacency(ctx, (uint32_t)n, rows.data(), deg.data(), (uint32_t)sizeof(Edge), (uint32_t)offsetof(Edge, to), 4u, dg);
}
}  // namespace

extern "C" {

// "Export once per sampled graph": the device graph of a handle's current adjacency lists.
int nech_upload_cuda(const void *p, nec_ctx *ctx, nec_graph **out) {
    if (!p || !ctx || !out) return NEC_EINVAL;
    return guarded(ctx, [&]() -> int {
        return with_graph(static_cast<const Handle *>(p), [&](auto &g) -> int { return upload_graph(ctx, g, out); }); });
}

// The drop-in call: <ensemble or graph>.vertex_load(include_endpoints).  out: n doubles.  Errors are reported
// through nec_last_error(ctx).
int nech_vertex_load_cuda(const void *p, nec_ctx *ctx, int include_endpoints, double *out) {
    if (!p || !ctx) return NEC_EINVAL;
    return guarded(ctx, [&]() -> int {
        return with_graph(static_cast<const Handle *>(p), [&](auto &g) -> int {
            nec_graph *dg = nullptr;
            int rc = upload_graph(ctx, g, &dg);
            if (rc != NEC_OK) return rc;
            rc = nec_vertex_load(ctx, dg, include_endpoints, out);
            nec_graph_free(ctx, dg);
            return rc;
        }); });
}

// SimpleSample::randomize of an ErEnsembleC ON THE DEVICE (nec_erc_randomize: Pcg64 jump-ahead, one draw per pair,
// bit-identical to er_c.rs:128-143), leaving the sampled graph device-resident for the measurement that follows
// (benches/bench_vertex_load.rs:20-26).  The host mirror is brought in step - adjacency lists, edge count and
// generator - from the device's CSR, so m_step / export / vertex_load on the handle keep working afterwards.
int nech_erc_randomize_cuda(void *p, nec_ctx *ctx, nec_graph **out) {
    if (!p || !ctx || !out) return NEC_EINVAL;
    return guarded(ctx, [&]() -> int {
        Handle *h = static_cast<Handle *>(p);
        if (h->kind != Kind::ErC) { nec_internal_set_error(ctx, "randomize on the device is implemented for ErEnsembleC"); return NEC_EINVAL; }
        ErEnsembleC &e = *h->erc;
        const uint32_t n = (uint32_t)e.graph.vertex_count();
        const uint64_t in[4] = {(uint64_t)e.rng.state, (uint64_t)(e.rng.state >> 64), (uint64_t)e.rng.increment, (uint64_t)(e.rng.increment >> 64)};
        uint64_t after[4];
        nec_graph *dg = nullptr;
        int rc = nec_erc_randomize(ctx, n, e.prob, in, after, &dg);
        if (rc != NEC_OK) return rc;
        uint64_t nnz = 0;
        nec_graph_dims(dg, nullptr, &nnz);
        std::vector<uint64_t> rp((size_t)n + 1);
        std::vector<uint32_t> col(nnz);
        rc = nec_graph_download(ctx, dg, rp.data(), col.data());
        if (rc != NEC_OK) { nec_graph_free(ctx, dg); return rc; }
        e.graph.clear_edges();
        for (uint32_t v = 0; v < n; ++v) {
            auto &adj = e.graph.vertices[v].adj;
            adj.resize(rp[v + 1] - rp[v]);
            for (uint64_t k = rp[v]; k < rp[v + 1]; ++k) adj[k - rp[v]] = PlainEdge{col[k]};
        }
        e.graph.edge_count = nnz / 2;
        e.rng.state = ((u128)after[1] << 64) | (u128)after[0];
        *out = dg;
        return NEC_OK; });
}

// Markov-chain shape (benches/common/mod.rs:79-90): `count` m_steps on every chain, then
// vertex_load(include_endpoints) of ALL chains' graphs (one CTA per graph).  out receives the load
// vectors back to back.  count == 0 measures without stepping.
//
// Stepping and export run on all host cores; nec_vertex_load_batch is itself a copy-in / kernel / copy-out
// pipeline.  (Stepping the next group of chains on a helper thread while the GPU works was tried and lost:
// two OpenMP teams on the same cores stall each other for longer than the 3 ms they would hide.)
int nech_chains_step_and_load_cuda(void *const *handles, uint32_t n_chains, uint64_t m_steps, nec_ctx *ctx,
                                   int include_endpoints, double *out) {
    if (!ctx || (n_chains && !handles)) return NEC_EINVAL;
    return guarded(ctx, [&]() -> int {
        const bool verbose = getenv("NEC_SMALL_VERBOSE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        g_chain_stats = nec_stats{};
        ChainGroup &grp = g_group;
        prepare_group(grp, handles, 0, n_chains, m_steps);
        if (grp.bad) {
            nec_internal_set_error(ctx, grp.bad == -2 ? "adjacency lists disagree with edge_count" : "chain has no Markov step");
            return NEC_EINVAL;
        }
        const double t1 = now();
        const int rc = nec_vertex_load_batch(ctx, n_chains, grp.n_per.data(), grp.rp_off.data(), grp.rp.data(),
                                             grp.col_off.data(), grp.col.data(), include_endpoints, out);
        if (rc == NEC_OK) nec_get_stats(ctx, &g_chain_stats);
        if (verbose) fprintf(stderr, "[nech chains] step+export %.2f ms, batch call %.2f ms\n", t1 - t0, now() - t1);
        return rc; });
}
// Counters of the last nech_chains_step_and_load_cuda call of this thread.
void nech_last_chain_stats(nec_stats *out) { *out = g_chain_stats; }

}  // extern "C"

// ---- device-resident chains: the graphs stay on the GPU, a step ships list operations ----------------------------
namespace {
struct DeviceChainBatch {
    nec_ctx *ctx = nullptr;
    std::vector<Handle *> handles;
    std::vector<std::vector<uint64_t>> logs;         // per chain: the list operations of the current step
    std::vector<uint64_t> seen_mutations, seen_wholesale;
    std::vector<uint32_t> op_off;
    std::vector<uint64_t> ops;
    nec_chains *dev = nullptr;
    uint32_t slack = 8;
    uint64_t h2d_last = 0, uploads = 0;
    nec_stats stats{};
};

template <class F>
auto with_graph_mut(Handle *h, F &&f) {
    if (h->kind == Kind::Sw) return f(h->sw->graph);
    return f(*const_cast<Graph *>(h->plain_graph()));
}

// (re-)create the device copy from the chains' current adjacency lists
int upload_chains(DeviceChainBatch &b) {
    if (b.dev) { nec_chains_free(b.ctx, b.dev); b.dev = nullptr; }
    ChainGroup &grp = g_group;
    prepare_group(grp, reinterpret_cast<void *const *>(b.handles.data()), 0, (uint32_t)b.handles.size(), 0);
    if (grp.bad) { nec_internal_set_error(b.ctx, "adjacency lists disagree with edge_count"); return NEC_EINVAL; }
    int rc = nec_chains_create(b.ctx, (uint32_t)b.handles.size(), grp.n_per.data(), grp.rp_off.data(), grp.rp.data(), grp.col_off.data(),
                               grp.col.data(), b.slack, &b.dev);
    if (rc != NEC_OK) return rc;
    for (size_t g = 0; g < b.handles.size(); ++g)
        with_graph_mut(b.handles[g], [&](auto &gr) { b.seen_mutations[g] = gr.mutations; b.seen_wholesale[g] = gr.wholesale_changes; return 0; });
    b.uploads++;
    b.h2d_last = (uint64_t)b.handles.size() * 20 + (grp.rp_off.empty() ? 0 : (grp.rp_off.back() + grp.n_per.back() + 1) * 4) + grp.col.size() * 4;
    return NEC_OK;
}
}  // namespace

extern "C" {

void *nech_chain_batch_new(void *const *handles, uint32_t n_chains, nec_ctx *ctx) {
    if (!ctx || !handles || !n_chains) return nullptr;
    DeviceChainBatch *b = nullptr;
    int rc = guarded(ctx, [&]() -> int {
        b = new DeviceChainBatch();
        b->ctx = ctx;
        b->handles.resize(n_chains);
        for (uint32_t g = 0; g < n_chains; ++g) b->handles[g] = static_cast<Handle *>(handles[g]);
        b->logs.resize(n_chains); b->seen_mutations.assign(n_chains, 0); b->seen_wholesale.assign(n_chains, 0);
        b->op_off.assign((size_t)n_chains + 1, 0);
        return upload_chains(*b); });
    if (rc != NEC_OK) { if (b) { if (b->dev) nec_chains_free(ctx, b->dev); delete b; } return nullptr; }
    return b;
}

void nech_chain_batch_free(void *batch) {
    DeviceChainBatch *b = static_cast<DeviceChainBatch *>(batch);
    if (!b) return;
    if (b->dev) nec_chains_free(b->ctx, b->dev);
    delete b;
}

// `m_steps` Markov steps on every chain (all host cores; the chains' own generators), the resulting list
// operations shipped to the device copy and replayed there, then vertex_load(include_endpoints) of all graphs.
// out == NULL: the loads stay in the pinned buffer nech_chain_batch_result returns.
int nech_chain_batch_step(void *batch, uint64_t m_steps, int include_endpoints, double *out) {
    DeviceChainBatch *b = static_cast<DeviceChainBatch *>(batch);
    if (!b) return NEC_EINVAL;
    return guarded(b->ctx, [&]() -> int {
        const size_t n = b->handles.size();
        int bad = 0, stale = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad, stale)
        for (int64_t g = 0; g < (int64_t)n; ++g) {
            Handle *h = b->handles[g];
            std::vector<uint64_t> &log = b->logs[g];
            log.clear();
            with_graph_mut(h, [&](auto &gr) {
                if (gr.mutations != b->seen_mutations[g] || gr.wholesale_changes != b->seen_wholesale[g]) stale += 1;   // stepped behind our back
                gr.oplog = &log;
                return 0;
            });
            for (uint64_t k = 0; k < m_steps; ++k)
                if (!h->m_step()) { bad += 1; break; }
            with_graph_mut(h, [&](auto &gr) {
                gr.oplog = nullptr;
                if (gr.wholesale_changes != b->seen_wholesale[g]) stale += 1;
                b->seen_mutations[g] = gr.mutations;
                return 0;
            });
        }
        if (bad) { nec_internal_set_error(b->ctx, "chain has no Markov step"); return NEC_EINVAL; }
        int rc;
        if (stale) {
            // some chain changed in a way the log does not cover (randomize, steps outside this batch): start over
            if ((rc = upload_chains(*b)) != NEC_OK) return rc;
            rc = nec_chains_step_and_load(b->ctx, b->dev, nullptr, nullptr, include_endpoints, out);
        } else {
            uint64_t total = 0;
            for (size_t g = 0; g < n; ++g) { b->op_off[g] = (uint32_t)total; total += b->logs[g].size(); }
            b->op_off[n] = (uint32_t)total;
            b->ops.resize(total);
#pragma omp parallel for schedule(static)
            for (int64_t g = 0; g < (int64_t)n; ++g)
                if (!b->logs[g].empty()) std::memcpy(b->ops.data() + b->op_off[g], b->logs[g].data(), b->logs[g].size() * 8);
            b->h2d_last = total * 8 + (total ? (n + 1) * 4 : 0);
            rc = nec_chains_step_and_load(b->ctx, b->dev, total ? b->op_off.data() : nullptr, b->ops.data(), include_endpoints, out);
            if (rc == NEC_EINTERNAL) {
                // an adjacency list outgrew its row: rebuild the device copy from the (already stepped) host lists
                b->slack *= 2;
                if ((rc = upload_chains(*b)) != NEC_OK) return rc;
                rc = nec_chains_step_and_load(b->ctx, b->dev, nullptr, nullptr, include_endpoints, out);
            }
        }
        if (rc == NEC_OK) nec_get_stats(b->ctx, &b->stats);
        return rc; });
}

const double *nech_chain_batch_result(const void *batch) {
    const DeviceChainBatch *b = static_cast<const DeviceChainBatch *>(batch);
    return b && b->dev ? nec_chains_result(b->dev) : nullptr;
}
void nech_chain_batch_stats(const void *batch, nec_stats *out, uint64_t *h2d_bytes_last_step, uint64_t *uploads) {
    const DeviceChainBatch *b = static_cast<const DeviceChainBatch *>(batch);
    if (out) *out = b->stats;
    if (h2d_bytes_last_step) *h2d_bytes_last_step = b->h2d_last;
    if (uploads) *uploads = b->uploads;
}
// the device copy's adjacency as a compact CSR batch (see nec_chains_download)
int nech_chain_batch_download(void *batch, uint32_t *row_ptr_concat, uint32_t *col_idx_concat, uint64_t col_capacity, uint64_t *nnz) {
    DeviceChainBatch *b = static_cast<DeviceChainBatch *>(batch);
    if (!b || !b->dev) return NEC_EINVAL;
    return nec_chains_download(b->ctx, b->dev, row_ptr_concat, col_idx_concat, col_capacity, nnz);
}

}  // extern "C"
