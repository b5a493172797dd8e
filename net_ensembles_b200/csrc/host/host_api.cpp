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
#include <type_traits>
#include <utility>
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

namespace {
// adjacency lists exactly as given, validated: ids < n, no self-loops, no repeated entry within a list, every entry (v, x)
// matched by an entry (x, v); E = PlainEdge, or SwEdge with the root target of every entry (NO_ROOT: loose) where exactly one
// of an edge's two entries must be the root
template <class E>
GenericGraph<E> lists_to_graph(uint64_t n, const uint64_t *row_ptr, const uint32_t *col_idx, const uint32_t *root_targets) {
    if (n > 0 && (!row_ptr || row_ptr[0] != 0)) throw std::runtime_error("adjacency lists: row_ptr must start at 0");
    GenericGraph<E> g(n);
    for (uint64_t v = 0; v < n; ++v) {
        if (row_ptr[v + 1] < row_ptr[v]) throw std::runtime_error("adjacency lists: row_ptr not monotone");
        auto &adj = g.vertices[v].adj;
        adj.reserve(row_ptr[v + 1] - row_ptr[v]);
        for (uint64_t e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
            const uint32_t x = col_idx[e];
            if (x >= n) throw std::runtime_error("adjacency lists: neighbour id out of range (vertex " + std::to_string(v) + ")");
            if (x == v) throw std::runtime_error("adjacency lists: self-loop at vertex " + std::to_string(v));
            if (g.vertices[v].is_adjacent(x)) throw std::runtime_error("adjacency lists: repeated entry in the list of vertex " + std::to_string(v));
            if constexpr (std::is_same<E, SwEdge>::value) {
                const uint32_t rt = root_targets ? root_targets[e] : NO_ROOT;
                if (rt != NO_ROOT && rt >= n) throw std::runtime_error("adjacency lists: root target out of range (vertex " + std::to_string(v) + ")");
                adj.push_back(SwEdge{x, rt});
            } else
                adj.push_back(PlainEdge{x});
        }
    }
    for (uint64_t v = 0; v < n; ++v)
        for (const auto &e : g.vertices[v].adj) {
            const long back = g.vertices[e.to].position_of((uint32_t)v);
            if (back < 0) throw std::runtime_error("adjacency lists: entry (" + std::to_string(v) + ", " + std::to_string(e.to) + ") has no counterpart");
            if constexpr (std::is_same<E, SwEdge>::value)
                if (e.is_root() == g.vertices[e.to].adj[(size_t)back].is_root())
                    throw std::runtime_error("adjacency lists: edge (" + std::to_string(v) + ", " + std::to_string(e.to) + ") must have exactly one root entry");
        }
    g.edge_count = (n ? row_ptr[n] : 0) / 2;
    return g;
}
std::vector<std::pair<uint32_t, uint32_t>> pairs_of(const uint32_t *flat, uint64_t count) {
    std::vector<std::pair<uint32_t, uint32_t>> v;
    v.reserve(count);
    for (uint64_t k = 0; k < count; ++k) v.emplace_back(flat[2 * k], flat[2 * k + 1]);
    return v;
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
// Resume an ensemble from a save-state of the crate (serde dump: er_c.rs:71-80, er_m.rs:157-170, sw.rs:190-200): adjacency lists in
// storage order, parameters, Pcg64 state and increment (4 x u64, low word first), for kind 2 (ErEnsembleM) its three pair lists
// flattened a0, b0, a1, b1, ..., for kind 3 (SwEnsemble) the root target of every adjacency entry (0xFFFFFFFF: loose entry).
// kind 1 = ErEnsembleC (p0 = c_target, p1 = prob), 2 = ErEnsembleM (m), 3 = SwEnsemble (p0 = r_prob).  Nothing is drawn.
void *nech_restore(int kind, uint64_t n, const uint64_t *row_ptr, const uint32_t *col_idx, const uint32_t *root_targets,
                   double p0, double p1, uint64_t m, const uint64_t *rng4,
                   const uint32_t *all_edges, uint64_t n_all, const uint32_t *possible_edges, uint64_t n_possible,
                   const uint32_t *current_edges, uint64_t n_current) {
    return guarded([&]() -> void * {
        if (!rng4) throw std::runtime_error("nech_restore: generator state missing");
        const Pcg64 rng = Pcg64::from_state((u128)rng4[0] | ((u128)rng4[1] << 64), (u128)rng4[2] | ((u128)rng4[3] << 64));
        if (kind == 1) {
            auto h = new Handle{Kind::ErC};
            std::unique_ptr<Handle> guard(h);
            h->erc.reset(new ErEnsembleC(lists_to_graph<PlainEdge>(n, row_ptr, col_idx, nullptr), p0, p1, rng));
            return guard.release();
        }
        if (kind == 2) {
            auto h = new Handle{Kind::ErM};
            std::unique_ptr<Handle> guard(h);
            h->erm.reset(new ErEnsembleM(lists_to_graph<PlainEdge>(n, row_ptr, col_idx, nullptr), (size_t)m, rng,
                                         pairs_of(all_edges, n_all), pairs_of(possible_edges, n_possible), pairs_of(current_edges, n_current)));
            return guard.release();
        }
        if (kind == 3) {
            auto h = new Handle{Kind::Sw};
            std::unique_ptr<Handle> guard(h);
            h->sw.reset(new SwEnsemble(lists_to_graph<SwEdge>(n, row_ptr, col_idx, root_targets), p0, rng));
            return guard.release();
        }
        throw std::runtime_error("nech_restore: kind must be 1 (ErEnsembleC), 2 (ErEnsembleM) or 3 (SwEnsemble)");
    }, nullptr);
}
// ErEnsembleM's pair lists (which: 0 all_edges, 1 possible_edges, 2 current_edges): size, and the pairs flattened
uint64_t nech_erm_list_size(const void *p, int which) {
    auto h = static_cast<const Handle *>(p);
    if (h->kind != Kind::ErM || which < 0 || which > 2) return 0;
    return which == 0 ? h->erm->all_edges.size() : which == 1 ? h->erm->possible_edges.size() : h->erm->current_edges.size();
}
int nech_erm_list(const void *p, int which, uint32_t *out_pairs) {
    auto h = static_cast<const Handle *>(p);
    if (h->kind != Kind::ErM || which < 0 || which > 2) { g_err = "nech_erm_list: not an ErEnsembleM"; return -1; }
    const auto &v = which == 0 ? h->erm->all_edges : which == 1 ? h->erm->possible_edges : h->erm->current_edges;
    for (size_t k = 0; k < v.size(); ++k) { out_pairs[2 * k] = v[k].first; out_pairs[2 * k + 1] = v[k].second; }
    return 0;
}

// A plain graph with EXACTLY these adjacency lists, in this order (the order defines the rounding of vertex_load and is
// what the crate's serde save-states store: graph.vertices[i].adj).  Validated: ids < n, no self-loops, no repeated entry
// within a list, every entry (v, x) matched by an entry (x, v).
void *nech_graph_from_adjacency(uint64_t n, const uint64_t *row_ptr, const uint32_t *col_idx) {
    return guarded([&]() -> void * {
        std::unique_ptr<Graph> g(new Graph(lists_to_graph<PlainEdge>(n, row_ptr, col_idx, nullptr)));
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
