// C entry points over the host-side graph/ensemble mirror (nec_graphs.hpp) for the Python
// harness (ctypes).  `nech_` = net_ensembles_cuda host.  These build the INPUT graphs of the
// hot path and call it through the public C ABI; they contain no load arithmetic and no CPU
// implementation of vertex_load.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../../include/net_ensembles_cuda.h"
#include "nec_graphs.hpp"

using namespace nec_host;

namespace {
thread_local std::string g_err;

enum class Kind { Plain, ErC, ErM, Sw, BA };

struct Handle {
    Kind kind;
    std::unique_ptr<Graph> plain;
    std::unique_ptr<ErEnsembleC> erc;
    std::unique_ptr<ErEnsembleM> erm;
    std::unique_ptr<SwEnsemble> sw;
    std::unique_ptr<BAensemble> ba;

    const Graph *plain_graph() const {
        switch (kind) {
            case Kind::Plain: return plain.get();
            case Kind::ErC: return &erc->graph;
            case Kind::ErM: return &erm->graph;
            case Kind::BA: return &ba->graph;
            default: return nullptr;
        }
    }
    Pcg64 *rng() {
        switch (kind) {
            case Kind::ErC: return &erc->rng;
            case Kind::ErM: return &erm->rng;
            case Kind::Sw: return &sw->rng;
            case Kind::BA: return &ba->rng;
            default: return nullptr;
        }
    }
};

template <class F>
auto guarded(F &&f, decltype(f()) on_error) -> decltype(f()) {
    try {
        return f();
    } catch (const std::exception &e) {
        g_err = e.what();
        return on_error;
    }
}

template <class F>
auto with_graph(const Handle *h, F &&f) {
    if (h->kind == Kind::Sw) return f(h->sw->graph);
    return f(*h->plain_graph());
}
}  // namespace

extern "C" {

const char *nech_last_error(void) { return g_err.c_str(); }

void *nech_graph_new(uint64_t n) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::Plain}; h->plain.reset(new Graph(n)); return h; }, nullptr);
}
void *nech_erc_new(uint64_t n, double c, uint64_t seed) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::ErC}; h->erc.reset(new ErEnsembleC(n, c, Pcg64::seed_from_u64(seed))); return h; }, nullptr);
}
void *nech_erm_new(uint64_t n, uint64_t m, uint64_t seed) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::ErM}; h->erm.reset(new ErEnsembleM(n, m, Pcg64::seed_from_u64(seed))); return h; }, nullptr);
}
void *nech_sw_new(uint64_t n, double r_prob, uint64_t seed, uint64_t ring_distance) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::Sw}; h->sw.reset(new SwEnsemble(n, r_prob, Pcg64::seed_from_u64(seed), ring_distance)); return h; }, nullptr);
}
void *nech_ba_new(uint64_t n, uint64_t m, uint64_t source_n, uint64_t seed) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::BA}; h->ba.reset(new BAensemble(n, Pcg64::seed_from_u64(seed), m, source_n)); return h; }, nullptr);
}
// scalable, distribution-equivalent producers for the large benchmark shapes (plain Graph)
void *nech_gnm_scalable(uint64_t n, uint64_t m, uint64_t seed) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::Plain}; h->plain.reset(new Graph(gnm_scalable(n, m, Pcg64::seed_from_u64(seed)))); return h; }, nullptr);
}
void *nech_ba_scalable(uint64_t n, uint64_t m, uint64_t source_n, uint64_t seed) {
    return guarded([&]() -> void * { auto h = new Handle{Kind::Plain}; h->plain.reset(new Graph(ba_scalable(n, m, source_n, Pcg64::seed_from_u64(seed)))); return h; }, nullptr);
}
void nech_free(void *p) { delete static_cast<Handle *>(p); }

uint64_t nech_vertex_count(const void *p) {
    return with_graph(static_cast<const Handle *>(p), [](auto &g) { return (uint64_t)g.vertex_count(); });
}
uint64_t nech_edge_count(const void *p) {
    return with_graph(static_cast<const Handle *>(p), [](auto &g) { return (uint64_t)g.edge_count; });
}
uint64_t nech_nnz(const void *p) {
    return with_graph(static_cast<const Handle *>(p), [](auto &g) { return (uint64_t)g.nnz(); });
}

// 1 = done, 0 = GraphErrors::EdgeExists / EdgeDoesNotExist, -1 = invalid indices
int nech_add_edge(void *p, uint32_t a, uint32_t b) {
    auto h = static_cast<Handle *>(p);
    return guarded([&]() -> int {
        if (h->kind == Kind::Sw) return h->sw->graph.add_edge(a, b) ? 1 : 0;
        return const_cast<Graph *>(h->plain_graph())->add_edge(a, b) ? 1 : 0; }, -1);
}
int nech_remove_edge(void *p, uint32_t a, uint32_t b) {
    auto h = static_cast<Handle *>(p);
    return guarded([&]() -> int {
        if (h->kind == Kind::Sw) return h->sw->graph.remove_edge(a, b) ? 1 : 0;
        return const_cast<Graph *>(h->plain_graph())->remove_edge(a, b) ? 1 : 0; }, -1);
}

// SimpleSample::randomize
int nech_randomize(void *p) {
    auto h = static_cast<Handle *>(p);
    return guarded([&]() -> int {
        switch (h->kind) {
            case Kind::ErC: h->erc->randomize(); return 0;
            case Kind::ErM: h->erm->randomize(); return 0;
            case Kind::Sw: h->sw->randomize(); return 0;
            case Kind::BA: h->ba->randomize(); return 0;
            default: g_err = "randomize: not an ensemble"; return -1;
        } }, -1);
}
// `count` Markov steps (MarkovChain::m_steps_quiet of the external `sampling` crate is a
// plain loop over m_step; BAensemble has no Markov chain in the reference)
int nech_m_steps(void *p, uint64_t count) {
    auto h = static_cast<Handle *>(p);
    return guarded([&]() -> int {
        for (uint64_t k = 0; k < count; ++k) switch (h->kind) {
            case Kind::ErC: h->erc->m_step(); break;
            case Kind::ErM: h->erm->m_step(); break;
            case Kind::Sw: h->sw->m_step(); break;
            default: g_err = "m_steps: ensemble has no Markov chain"; return -1;
        }
        return 0; }, -1);
}

int nech_export_csr(const void *p, uint64_t *row_ptr, uint32_t *col_idx) {
    return guarded([&]() -> int {
        with_graph(static_cast<const Handle *>(p), [&](auto &g) { g.export_csr(row_ptr, col_idx); return 0; });
        return 0; }, -1);
}
// per adjacency entry (CSR order): pristine ring target of a root edge, 0xFFFFFFFF otherwise
int nech_sw_root_targets(const void *p, uint32_t *out) {
    auto h = static_cast<const Handle *>(p);
    if (h->kind != Kind::Sw) { g_err = "not a small-world ensemble"; return -1; }
    size_t at = 0;
    for (auto &c : h->sw->graph.vertices)
        for (auto &e : c.adj) out[at++] = e.root_target;
    return 0;
}
// generator state as 4 x u64: state lo, state hi, increment lo, increment hi
int nech_rng_state(void *p, uint64_t *out4) {
    Pcg64 *r = static_cast<Handle *>(p)->rng();
    if (!r) { g_err = "no rng"; return -1; }
    out4[0] = (uint64_t)r->state; out4[1] = (uint64_t)(r->state >> 64);
    out4[2] = (uint64_t)r->increment; out4[3] = (uint64_t)(r->increment >> 64);
    return 0;
}

// The drop-in call: <ensemble or graph>.vertex_load(include_endpoints) — export the
// adjacency lists to CSR once, upload, run the CUDA path, free.  out: n doubles.
int nech_vertex_load_cuda(const void *p, nec_ctx *ctx, int include_endpoints, double *out) {
    return guarded([&]() -> int {
        return with_graph(static_cast<const Handle *>(p), [&](auto &g) -> int {
            std::vector<uint64_t> row_ptr(g.vertex_count() + 1);
            std::vector<uint32_t> col(g.nnz());
            g.export_csr(row_ptr.data(), col.data());
            nec_graph *dg = nullptr;
            int rc = nec_graph_upload(ctx, (uint32_t)g.vertex_count(), col.size(), row_ptr.data(), col.data(), &dg);
            if (rc != NEC_OK) { g_err = nec_last_error(ctx); return rc; }
            rc = nec_vertex_load(ctx, dg, include_endpoints, out);
            if (rc != NEC_OK) g_err = nec_last_error(ctx);
            nec_graph_free(ctx, dg);
            return rc;
        }); }, NEC_EINVAL);
}

// Markov-chain shape (benches/common/mod.rs:79-90): `count` m_steps on every chain, then
// vertex_load(include_endpoints) of ALL chains' graphs (one CTA per graph).  out receives the load
// vectors back to back.  count == 0 measures without stepping.
//
// Stepping and export run on all host cores; nec_vertex_load_batch is itself a copy-in / kernel / copy-out
// pipeline.  (Stepping the next group of chains on a helper thread while the GPU works was tried and lost:
// two OpenMP teams on the same cores stall each other for longer than the 3 ms they would hide.)
namespace {
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
        for (uint64_t k = 0; k < m_steps; ++k) switch (h->kind) {
            case Kind::ErC: h->erc->m_step(); break;
            case Kind::ErM: h->erm->m_step(); break;
            case Kind::Sw: h->sw->m_step(); break;
            default: bad += 1; k = m_steps; break;
        }
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

int nech_chains_step_and_load_cuda(void *const *handles, uint32_t n_chains, uint64_t m_steps, nec_ctx *ctx,
                                   int include_endpoints, double *out) {
    return guarded([&]() -> int {
        const bool verbose = getenv("NEC_SMALL_VERBOSE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        g_chain_stats = nec_stats{};
        ChainGroup &grp = g_group;
        prepare_group(grp, handles, 0, n_chains, m_steps);
        if (grp.bad) { g_err = grp.bad == -2 ? "adjacency lists disagree with edge_count" : "chain has no Markov step"; return NEC_EINVAL; }
        const double t1 = now();
        const int rc = nec_vertex_load_batch(ctx, n_chains, grp.n_per.data(), grp.rp_off.data(), grp.rp.data(),
                                             grp.col_off.data(), grp.col.data(), include_endpoints, out);
        if (rc != NEC_OK) g_err = nec_last_error(ctx);
        else nec_get_stats(ctx, &g_chain_stats);
        if (verbose) fprintf(stderr, "[nech chains] step+export %.2f ms, batch call %.2f ms\n", t1 - t0, now() - t1);
        return rc; }, NEC_EINVAL);
}
// Counters of the last nech_chains_step_and_load_cuda call of this thread.
void nech_last_chain_stats(nec_stats *out) { *out = g_chain_stats; }

// raw generator access, for the known-answer tests of the rand restatement
void nech_pcg64_draws(uint64_t seed, uint64_t count, uint64_t *out_u64) {
    Pcg64 r = Pcg64::seed_from_u64(seed);
    for (uint64_t k = 0; k < count; ++k) out_u64[k] = r.next_u64();
}

}  // extern "C"
