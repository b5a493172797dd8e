// C entry points over the host-side graph/ensemble mirror (nec_graphs.hpp) for the Python
// harness (ctypes): libnec_host.so.  `nech_` = net_ensembles_cuda host.  These build and step the
// INPUT graphs of the hot path and export them as CSR; they contain no load arithmetic, no CPU
// implementation of vertex_load and no CUDA (the library links neither the CUDA runtime nor
// libnet_ensembles_cuda.so, so harness code that only needs graphs - e.g. the CPU baseline arm of
// bench.py - never maps the product library).  The drop-in call over a handle lives in
// host_bridge.cpp, inside the CUDA library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "nec_graphs.hpp"
#include "nec_handle.hpp"

using namespace nec_host;

namespace {
thread_local std::string g_err;

template <class F>
auto guarded(F &&f, decltype(f()) on_error) -> decltype(f()) {
    try {
        return f();
    } catch (const std::exception &e) {
        g_err = e.what();
        return on_error;
    }
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
// A plain graph with EXACTLY these adjacency lists, in this order (the order defines the rounding of vertex_load and is
// what the crate's serde save-states store: graph.vertices[i].adj).  Validated: ids < n, no self-loops, no repeated entry
// within a list, every entry (v, x) matched by an entry (x, v).
void *nech_graph_from_adjacency(uint64_t n, const uint64_t *row_ptr, const uint32_t *col_idx) {
    return guarded([&]() -> void * {
        if (n > 0 && (!row_ptr || row_ptr[0] != 0)) throw std::runtime_error("nech_graph_from_adjacency: row_ptr must start at 0");
        std::unique_ptr<Graph> g(new Graph(n));
        for (uint64_t v = 0; v < n; ++v) {
            if (row_ptr[v + 1] < row_ptr[v]) throw std::runtime_error("nech_graph_from_adjacency: row_ptr not monotone");
            auto &adj = g->vertices[v].adj;
            adj.reserve(row_ptr[v + 1] - row_ptr[v]);
            for (uint64_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                const uint32_t x = col_idx[e];
                if (x >= n) throw std::runtime_error("nech_graph_from_adjacency: neighbour id out of range (vertex " + std::to_string(v) + ")");
                if (x == v) throw std::runtime_error("nech_graph_from_adjacency: self-loop at vertex " + std::to_string(v));
                if (g->vertices[v].is_adjacent(x)) throw std::runtime_error("nech_graph_from_adjacency: repeated entry in the list of vertex " + std::to_string(v));
                adj.push_back(PlainEdge{x});
            }
        }
        for (uint64_t v = 0; v < n; ++v)
            for (const auto &e : g->vertices[v].adj)
                if (!g->vertices[e.to].is_adjacent((uint32_t)v))
                    throw std::runtime_error("nech_graph_from_adjacency: entry (" + std::to_string(v) + ", " + std::to_string(e.to) + ") has no counterpart");
        g->edge_count = (n ? row_ptr[n] : 0) / 2;
        auto h = new Handle{Kind::Plain};
        h->plain = std::move(g);
        return h;
    }, nullptr);
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

// raw generator access, for the known-answer tests of the rand restatement
void nech_pcg64_draws(uint64_t seed, uint64_t count, uint64_t *out_u64) {
    Pcg64 r = Pcg64::seed_from_u64(seed);
    for (uint64_t k = 0; k < count; ++k) out_u64[k] = r.next_u64();
}

}  // extern "C"
