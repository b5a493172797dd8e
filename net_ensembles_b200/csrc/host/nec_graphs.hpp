// Host-side mirror of the reference's graph containers and ensembles — the INPUT side of
// the vertex_load hot path (they produce the adjacency lists the exporter flattens).
// Nothing here is accelerated; it exists so that "identical seeded graphs" can be built
// without a Rust toolchain, and so the C++/Python host API reads like the crate's.
//
// Reference types mirrored (behaviour, not code):
//   GenericGraph{edge_count, vertices}   src/generic_graph/generic_graph.rs:33-58, add_edge :487, remove_edge :503
//   NodeContainer (Vec<usize> adj)       src/graph.rs:33-37, push :101-110, remove :114-142
//   SwEdge / SwContainer                 src/sw_graph.rs:16-19, :158-162, push :259-279, rewire :350-392, reset :398-442
//   SwGraph::reset_edge / rewire_edge    src/sw_graph.rs:453-483, init_ring generic_graph.rs:151-169
//   ErEnsembleC                          src/er_c.rs:128-174, :239-250, :349-358
//   ErEnsembleM                          src/er_m.rs:157-176, :202-236, :245-278
//   SwEnsemble                           src/sw.rs:214-225, :289-309, :325-343, :463-504
//   BAensemble                           src/barabasi_albert.rs:89-103, :199-229
// Adjacency ORDER is part of the contract: vertex_load's rounding depends on it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <unordered_set>
#include <vector>

#include "nec_rand.hpp"

namespace nec_host {

constexpr uint32_t NO_ROOT = 0xFFFFFFFFu;

// Adjacency mutations as a replayable log (device-resident Markov chains: the GPU copy of a chain's adjacency
// lists is patched with exactly the list operations the host performed, so ORDER stays identical).  Three
// primitives express every edge move of the crate: push x onto v's list (graph.rs:101-110, sw_graph.rs:259-279),
// swap_remove the entry x from v's list (graph.rs:134-142, sw_graph.rs:350-442) and retarget the entry x of v's
// list to y in place (rewire / reset of a small-world root edge).  One op = kind << 60 | v << 40 | x << 20 | y.
enum : uint64_t { OP_PUSH = 1, OP_SWAP_REMOVE = 2, OP_RETARGET = 3 };
constexpr uint32_t OP_MAX_VERTEX = (1u << 20) - 1;
inline uint64_t make_op(uint64_t kind, uint32_t v, uint32_t x, uint32_t y = 0) {
    return (kind << 60) | ((uint64_t)v << 40) | ((uint64_t)x << 20) | (uint64_t)y;
}

// ---- adjacency containers -------------------------------------------------------

struct PlainEdge {          // NodeContainer stores bare neighbour ids
    uint32_t to;
};
struct SwEdge {             // `to` + where a root edge pointed in the pristine ring
    uint32_t to;
    uint32_t root_target;   // NO_ROOT for a loose (non-root) entry
    bool is_root() const { return root_target != NO_ROOT; }
};

template <class E>
struct Container {
    std::vector<E> adj;
    long position_of(uint32_t other) const {
        for (size_t k = 0; k < adj.size(); ++k)
            if (adj[k].to == other) return (long)k;
        return -1;
    }
    bool is_adjacent(uint32_t other) const { return position_of(other) >= 0; }
    void swap_remove_at(size_t k) {
        adj[k] = adj.back();
        adj.pop_back();
    }
};

inline PlainEdge make_root_entry(PlainEdge *, uint32_t to) { return PlainEdge{to}; }
inline PlainEdge make_loose_entry(PlainEdge *, uint32_t to) { return PlainEdge{to}; }
inline SwEdge make_root_entry(SwEdge *, uint32_t to) { return SwEdge{to, to}; }
inline SwEdge make_loose_entry(SwEdge *, uint32_t to) { return SwEdge{to, NO_ROOT}; }

// ---- generic graph --------------------------------------------------------------

template <class E>
class GenericGraph {
public:
    std::vector<Container<E>> vertices;
    uint64_t edge_count = 0;
    std::vector<uint64_t> *oplog = nullptr;   // when set, every list mutation is appended (see make_op)
    uint64_t wholesale_changes = 0;           // bumped by mutations the log cannot express (clear, bulk rebuild)

    explicit GenericGraph(size_t n = 0) : vertices(n) {}
    uint64_t mutations = 0;                   // every list operation, logged or not (detects steps taken behind a log's back)
    void log(uint64_t kind, uint32_t v, uint32_t x, uint32_t y = 0) {
        ++mutations;
        if (oplog) oplog->push_back(make_op(kind, v, x, y));
    }
    size_t vertex_count() const { return vertices.size(); }
    size_t degree(size_t i) const { return vertices[i].adj.size(); }

    void clear_edges() {
        for (auto &c : vertices) c.adj.clear();
        edge_count = 0;
        ++wholesale_changes;
    }

    // false <=> GraphErrors::EdgeExists.  The edge is rooted at a (matters for SwEdge).
    bool add_edge(uint32_t a, uint32_t b) {
        check_pair(a, b);
        if (vertices[a].is_adjacent(b)) return false;
        vertices[a].adj.push_back(make_root_entry((E *)nullptr, b));
        vertices[b].adj.push_back(make_loose_entry((E *)nullptr, a));
        ++edge_count;
        log(OP_PUSH, a, b); log(OP_PUSH, b, a);
        return true;
    }
    // false <=> GraphErrors::EdgeDoesNotExist.  swap_remove on both lists: order changes.
    bool remove_edge(uint32_t a, uint32_t b) {
        check_pair(a, b);
        long pa = vertices[a].position_of(b);
        if (pa < 0) return false;
        vertices[a].swap_remove_at((size_t)pa);
        vertices[b].swap_remove_at((size_t)vertices[b].position_of(a));
        --edge_count;
        log(OP_SWAP_REMOVE, a, b); log(OP_SWAP_REMOVE, b, a);
        return true;
    }

    // every vertex linked to its next `distance` ring successors, rooted at the lower index
    void init_ring(size_t distance) {
        clear_edges();
        const size_t n = vertex_count();
        for (size_t i = 0; i < n; ++i)
            for (size_t step = 1; step <= distance; ++step) {
                size_t j = i + step;
                if (j >= n) j -= n;
                if (!add_edge((uint32_t)i, (uint32_t)j)) throw std::runtime_error("init_ring: edge exists");
            }
    }

    static GenericGraph complete(size_t n) {
        GenericGraph g(n);
        for (size_t i = 0; i < n; ++i) {
            g.vertices[i].adj.reserve(n - 1);
            for (size_t j = 0; j < n; ++j)
                if (j != i) g.vertices[i].adj.push_back(make_loose_entry((E *)nullptr, (uint32_t)j));
        }
        g.edge_count = (uint64_t)n * (n - 1) / 2;
        return g;
    }

    uint64_t nnz() const {
        uint64_t s = 0;
        for (auto &c : vertices) s += c.adj.size();
        return s;
    }
    // Order-preserving CSR: the flattening of container(i).neighbors() for i in 0..n.
    void export_csr(uint64_t *row_ptr, uint32_t *col_idx) const {
        uint64_t at = 0;
        for (size_t i = 0; i < vertices.size(); ++i) {
            row_ptr[i] = at;
            for (auto &e : vertices[i].adj) col_idx[at++] = e.to;
        }
        row_ptr[vertices.size()] = at;
    }

private:
    void check_pair(uint32_t a, uint32_t b) const {
        if (a >= vertices.size() || b >= vertices.size()) throw std::out_of_range("vertex index out of range");
        if (a == b) throw std::invalid_argument("self loops are not allowed");
    }
};

using Graph = GenericGraph<PlainEdge>;
using SwGraph = GenericGraph<SwEdge>;

// ---- small-world edge moves -------------------------------------------------------

enum class SwChange { Nothing, Rewire, Reset, Blocked, InvalidAdjacency, EdgeMissing };

// move root edge (i0 -> i1) so that it points to i2
inline SwChange sw_rewire_edge(SwGraph &g, uint32_t i0, uint32_t i1, uint32_t i2) {
    if (i1 == i2) return SwChange::Nothing;
    auto &c0 = g.vertices[i0];
    long p = c0.position_of(i1);
    if (p < 0) return SwChange::InvalidAdjacency;
    if (c0.is_adjacent(i2)) return SwChange::Blocked;
    auto &c1 = g.vertices[i1];
    c1.swap_remove_at((size_t)c1.position_of(i0));
    c0.adj[(size_t)p].to = i2;
    g.vertices[i2].adj.push_back(SwEdge{i0, NO_ROOT});
    g.log(OP_SWAP_REMOVE, i1, i0); g.log(OP_RETARGET, i0, i1, i2); g.log(OP_PUSH, i2, i0);
    return SwChange::Rewire;
}

// put root edge (i0 -> i1) back to its pristine ring target
inline SwChange sw_reset_edge(SwGraph &g, uint32_t i0, uint32_t i1) {
    auto &c0 = g.vertices[i0];
    long p0 = c0.position_of(i1);
    if (p0 < 0) return SwChange::EdgeMissing;
    auto &c1 = g.vertices[i1];
    long p1 = c1.position_of(i0);
    SwEdge &edge = c0.adj[(size_t)p0];
    if (!edge.is_root()) return SwChange::InvalidAdjacency;  // the crate panics here (debug_assert + unwrap)
    if (edge.root_target == edge.to) return SwChange::Nothing;
    uint32_t home = edge.root_target;
    if (c0.is_adjacent(home)) return SwChange::Blocked;
    edge.to = home;
    c1.swap_remove_at((size_t)p1);
    g.vertices[home].adj.push_back(SwEdge{i0, NO_ROOT});
    g.log(OP_RETARGET, i0, i1, home); g.log(OP_SWAP_REMOVE, i1, i0); g.log(OP_PUSH, home, i0);
    return SwChange::Reset;
}

// ---- ensembles ----------------------------------------------------------------------

struct ErEnsembleC {
    Graph graph;
    double c_target, prob;
    Pcg64 rng;
    ErEnsembleC(size_t n, double c, Pcg64 r) : graph(n), c_target(c), prob(c / (double)(n - 1)), rng(r) { randomize(); }
    // resume from a save-state (serde dump of the crate: graph, prob, c_target, rng): nothing is drawn
    ErEnsembleC(Graph &&g, double c, double p, Pcg64 r) : graph(std::move(g)), c_target(c), prob(p), rng(r) {}
    void randomize() {  // one uniform draw per unordered pair, lexicographic
        graph.clear_edges();
        const size_t n = graph.vertex_count();
        for (size_t i = 0; i < n; ++i)
            for (size_t j = i + 1; j < n; ++j)
                if (rng.gen_f64() <= prob) {   // (after clear_edges: a wholesale change, not logged entry by entry)
                    graph.vertices[i].adj.push_back(PlainEdge{(uint32_t)j});
                    graph.vertices[j].adj.push_back(PlainEdge{(uint32_t)i});
                    ++graph.edge_count;
                }
    }
    // one Markov step: pick an ordered pair of distinct vertices, then try to add
    // (w.p. prob) or remove the edge.  Returns +1 added, -1 removed, 0 nothing.
    int m_step() {
        const uint64_t n = graph.vertex_count();
        uint64_t a = rng.gen_below_u64(n);
        uint64_t b = rng.gen_below_u64(n - 1);
        if (b >= a) ++b;
        if (rng.gen_f64() <= prob) return graph.add_edge((uint32_t)a, (uint32_t)b) ? 1 : 0;
        return graph.remove_edge((uint32_t)a, (uint32_t)b) ? -1 : 0;
    }
};

struct ErEnsembleM {
    using Pair = std::pair<uint32_t, uint32_t>;
    Graph graph;
    size_t m;
    Pcg64 rng;
    std::vector<Pair> all_edges, possible_edges, current_edges;
    ErEnsembleM(size_t n, size_t m_, Pcg64 r) : graph(n), m(m_), rng(r) {
        const size_t total = n * (n - 1) / 2;  // O(n^2) memory, as in the reference
        if (m > total) throw std::invalid_argument("ErEnsembleM: more edges requested than a complete graph has");
        all_edges.reserve(total);
        for (size_t i = 0; i < n; ++i)
            for (size_t j = i + 1; j < n; ++j) all_edges.emplace_back((uint32_t)i, (uint32_t)j);
        randomize();
    }
    // resume from a save-state (graph, m, rng and the three pair lists, er_m.rs:157-170)
    ErEnsembleM(Graph &&g, size_t m_, Pcg64 r, std::vector<Pair> all, std::vector<Pair> possible, std::vector<Pair> current)
        : graph(std::move(g)), m(m_), rng(r), all_edges(std::move(all)), possible_edges(std::move(possible)), current_edges(std::move(current)) {
        const size_t n = graph.vertex_count(), total = n * (n - 1) / 2;
        if (all_edges.size() != total || possible_edges.size() + current_edges.size() != total || current_edges.size() != m || graph.edge_count != m)
            throw std::invalid_argument("ErEnsembleM: the save-state's edge lists do not fit n and m");
        for (auto &e : current_edges)
            if (e.first >= n || e.second >= n || !graph.vertices[e.first].is_adjacent(e.second))
                throw std::invalid_argument("ErEnsembleM: current_edges names an edge the graph does not have");
    }
    void randomize() {
        graph.clear_edges();
        shuffle(all_edges, rng);
        current_edges.assign(all_edges.begin(), all_edges.begin() + (long)m);
        for (auto &e : current_edges)
            if (!graph.add_edge(e.first, e.second)) throw std::runtime_error("ErEnsembleM: duplicate edge");
        possible_edges.assign(all_edges.begin() + (long)m, all_edges.end());
    }
    void m_step() {  // swap one present edge with one absent edge
        size_t ic = (size_t)rng.gen_below_u64(current_edges.size());
        size_t ip = (size_t)rng.gen_below_u64(possible_edges.size());
        Pair gone = current_edges[ic], born = possible_edges[ip];
        graph.remove_edge(gone.first, gone.second);
        graph.add_edge(born.first, born.second);
        std::swap(current_edges[ic], possible_edges[ip]);
    }
};

struct SwEnsemble {
    SwGraph graph;
    double r_prob;
    Pcg64 rng;
    // SwEnsemble::new (ring distance 2).  new_with_distance(d) builds a wider ring but the
    // randomisation still touches the first two root edges of every vertex, as the crate does.
    SwEnsemble(size_t n, double p, Pcg64 r, size_t ring_distance = 2) : graph(n), r_prob(p), rng(r) {
        graph.init_ring(ring_distance);
        randomize();
    }
    // resume from a save-state (graph with its root edges, r_prob, rng); the Markov step assumes two root edges per vertex
    SwEnsemble(SwGraph &&g, double p, Pcg64 r) : graph(std::move(g)), r_prob(p), rng(r) {
        for (size_t v = 0; v < graph.vertex_count(); ++v) {
            size_t roots = 0;
            for (auto &e : graph.vertices[v].adj) roots += e.is_root() ? 1 : 0;
            if (roots != 2) throw std::invalid_argument("SwEnsemble: a vertex of the save-state does not have two root edges");
        }
    }
    SwChange randomize_edge(uint32_t i0, uint32_t i1) {
        if (rng.gen_f64() <= r_prob) {
            uint64_t t = rng.gen_below_u64(graph.vertex_count() - 1);
            if (t >= i0) ++t;
            return sw_rewire_edge(graph, i0, i1, (uint32_t)t);
        }
        return sw_reset_edge(graph, i0, i1);
    }
    void randomize() {
        const size_t n = graph.vertex_count();
        for (size_t i = 0; i < n; ++i) {
            uint32_t target[2];
            int found = 0;
            for (auto &e : graph.vertices[i].adj)
                if (e.is_root() && found < 2) target[found++] = e.to;
            if (found < 2) throw std::runtime_error("SwEnsemble: vertex without two root edges");
            randomize_edge((uint32_t)i, target[0]);
            randomize_edge((uint32_t)i, target[1]);
        }
    }
    SwChange m_step() {
        uint64_t pick = rng.gen_below_u64(graph.edge_count);
        uint32_t v = (uint32_t)(pick / 2);
        int which = (int)(pick % 2), seen = 0;
        for (auto &e : graph.vertices[v].adj)
            if (e.is_root() && seen++ == which) return randomize_edge(v, e.to);
        throw std::runtime_error("SwEnsemble::m_step: root edge not found");
    }
};

struct BAensemble {
    Graph graph, source_graph;
    size_t m;
    Pcg64 rng;
    std::vector<uint64_t> weights;
    BAensemble(size_t n, Pcg64 r, size_t m_, size_t source_n) : graph(n), source_graph(Graph::complete(source_n)), m(m_), rng(r), weights(n, 0) {
        if (source_n < 2 || n <= source_n) throw std::invalid_argument("BAensemble: need 2 <= source_n < n");
        randomize();
    }
    // O(n^2): a fresh weighted sampler per arriving vertex, exactly like the crate.
    void randomize() {
        const size_t n = graph.vertex_count(), s = source_graph.vertex_count();
        graph.clear_edges();
        for (size_t i = 0; i < s; ++i) graph.vertices[i].adj = source_graph.vertices[i].adj;
        graph.edge_count = source_graph.edge_count;
        std::vector<uint32_t> order;
        for (size_t i = s; i < n; ++i) order.push_back((uint32_t)i);
        shuffle(order, rng);
        for (size_t i = 0; i < n; ++i) weights[i] = i < s ? source_graph.degree(i) : 0;
        for (uint32_t i : order) {
            WeightedIndex pick(weights);
            while (graph.degree(i) < m) graph.add_edge(i, (uint32_t)pick.sample(rng));
            weights[i] = graph.degree(i);
            for (auto &e : graph.vertices[i].adj) weights[e.to] += 1;
        }
    }
};

// ---- scalable, distribution-equivalent generators for the large benchmark shapes -----
// The reference constructors are O(n^2) in time or memory (er_m.rs:261-266,
// barabasi_albert.rs:214-215) and cannot reach N = 2^20 / 2^22.  These produce graphs from
// the same distributions with the same adjacency-list conventions, but are NOT
// bit-identical to what the crate would sample for a given seed.

// G(n,m): m distinct unordered pairs, uniformly; edges appended in acceptance order.
inline Graph gnm_scalable(size_t n, size_t m, Pcg64 rng) {
    Graph g(n);
    std::unordered_set<uint64_t> seen;
    seen.reserve(m * 2);
    while (g.edge_count < m) {
        uint64_t a = rng.gen_below_u64(n), b = rng.gen_below_u64(n - 1);
        if (b >= a) ++b;
        uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
        if (!seen.insert(lo * (uint64_t)n + hi).second) continue;
        g.vertices[a].adj.push_back(PlainEdge{(uint32_t)b});
        g.vertices[b].adj.push_back(PlainEdge{(uint32_t)a});
        ++g.edge_count;
    }
    return g;
}

// Barabasi-Albert: complete source graph, shuffled arrival order, each arrival links to m
// distinct existing vertices with probability proportional to degree (weights frozen while
// one vertex arrives, as in the crate) via a repeated-endpoint urn.
inline Graph ba_scalable(size_t n, size_t m, size_t source_n, Pcg64 rng) {
    Graph g(n);
    {
        Graph src = Graph::complete(source_n);
        for (size_t i = 0; i < source_n; ++i) g.vertices[i].adj = src.vertices[i].adj;
        g.edge_count = src.edge_count;
    }
    std::vector<uint32_t> urn;
    urn.reserve(2 * (g.edge_count + (n - source_n) * m));
    for (size_t i = 0; i < source_n; ++i)
        for (size_t k = 0; k < g.degree(i); ++k) urn.push_back((uint32_t)i);
    std::vector<uint32_t> order;
    for (size_t i = source_n; i < n; ++i) order.push_back((uint32_t)i);
    shuffle(order, rng);
    for (uint32_t i : order) {
        const size_t frozen = urn.size();
        while (g.degree(i) < m) {
            uint32_t t = urn[(size_t)rng.gen_below_u64(frozen)];
            g.add_edge(i, t);  // duplicate targets are simply redrawn
        }
        for (auto &e : g.vertices[i].adj) {
            urn.push_back(i);
            urn.push_back(e.to);
        }
    }
    return g;
}

}  // namespace nec_host
