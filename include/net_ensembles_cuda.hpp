// C++ mirror of the reference's operator interface for the vertex_load path, written over
// the C ABI only (include/net_ensembles_cuda.h).  The reference is compiled (Rust) code and
// no Rust toolchain exists in this environment, so the host side above the C ABI is C++;
// rust/net_ensembles_cuda.rs holds the equivalent Rust module a maintainer would add.
//
// Reference interface mirrored:
//   trait MeasurableGraphQuantities { fn vertex_load(&self, include_endpoints: bool) -> Vec<f64> }
//       src/traits/graph_traits.rs:267, blanket impl for E: AsRef<GenericGraph<T,A>> :272-331
//   GenericGraph::vertex_load                     src/generic_graph/generic_graph.rs:1310
// Anything that exposes `graph` (a nec_host::GenericGraph) or is one gets
// vertex_load_cuda(x, include_endpoints): same name, same argument meaning, returns a fresh
// vector owned by the caller.  The reference call is infallible; here a failing device call
// throws (the Rust shim panics) — there is no CPU fallback.
#pragma once
#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "net_ensembles_cuda.h"
#include "../net_ensembles_b200/csrc/host/nec_graphs.hpp"

namespace net_ensembles_cuda {

class Context {
public:
    explicit Context(int device = -1) {
        int dev = device;
        if (nec_init(&ctx_, device >= 0 ? &dev : nullptr, 1) != NEC_OK)
            throw std::runtime_error(std::string("nec_init: ") + nec_last_error(nullptr));
    }
    ~Context() { nec_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    nec_ctx *raw() const { return ctx_; }

private:
    nec_ctx *ctx_ = nullptr;
};

// Device-resident CSR of one sampled graph ("export once per sampled graph").
class DeviceGraph {
public:
    template <class E>
    DeviceGraph(Context &ctx, const nec_host::GenericGraph<E> &g) : ctx_(ctx), n_((uint32_t)g.vertex_count()) {
        // the containers keep a vertex's neighbours in one contiguous slice (AdjList::edges()): hand the slices
        // over and let the library flatten them into its pinned staging area (export fast path)
        std::vector<const void *> rows(n_);
        std::vector<uint64_t> deg(n_);
        for (uint32_t v = 0; v < n_; ++v) { rows[v] = g.vertices[v].adj.data(); deg[v] = g.vertices[v].adj.size(); }
        if (nec_graph_upload_adjacency(ctx.raw(), n_, rows.data(), deg.data(), (uint32_t)sizeof(E), (uint32_t)offsetof(E, to), 4u, &g_) != NEC_OK)
            throw std::runtime_error(std::string("nec_graph_upload_adjacency: ") + nec_last_error(ctx.raw()));
    }
    ~DeviceGraph() { nec_graph_free(ctx_.raw(), g_); }
    DeviceGraph(const DeviceGraph &) = delete;
    DeviceGraph &operator=(const DeviceGraph &) = delete;

    std::vector<double> vertex_load(bool include_endpoints) const {
        std::vector<double> out(n_);
        if (nec_vertex_load(ctx_.raw(), g_, include_endpoints ? 1 : 0, out.data()) != NEC_OK)
            throw std::runtime_error(std::string("nec_vertex_load: ") + nec_last_error(ctx_.raw()));
        return out;
    }

    // GenericGraph::diameter (generic_graph.rs:1116-1136): nullopt if empty or not connected
    std::optional<size_t> diameter() const {
        if (n_ == 0) return std::nullopt;
        std::vector<uint32_t> src(n_), ecc(n_), reached(n_);
        for (uint32_t v = 0; v < n_; ++v) src[v] = v;
        if (nec_distance_measures(ctx_.raw(), g_, src.data(), n_, ecc.data(), nullptr, reached.data()) != NEC_OK)
            throw std::runtime_error(std::string("nec_distance_measures: ") + nec_last_error(ctx_.raw()));
        if (reached[0] != n_) return std::nullopt;
        uint32_t best = 0;
        for (uint32_t v = 1; v < n_; ++v) best = ecc[v] > best ? ecc[v] : best;
        return best;
    }
    // GenericGraph::closeness_centrality (generic_graph.rs:1387-1403)
    std::vector<double> closeness_centrality() const {
        std::vector<uint32_t> src(n_);
        std::vector<uint64_t> sum(n_);
        for (uint32_t v = 0; v < n_; ++v) src[v] = v;
        if (nec_distance_measures(ctx_.raw(), g_, src.data(), n_, nullptr, sum.data(), nullptr) != NEC_OK)
            throw std::runtime_error(std::string("nec_distance_measures: ") + nec_last_error(ctx_.raw()));
        std::vector<double> out(n_);
        for (uint32_t v = 0; v < n_; ++v) out[v] = (double)(n_ - 1) / (double)sum[v];
        return out;
    }

private:
    Context &ctx_;
    nec_graph *g_ = nullptr;
    uint32_t n_;
};

// graph-like: a GenericGraph itself ...
template <class E>
const nec_host::GenericGraph<E> &as_graph(const nec_host::GenericGraph<E> &g) { return g; }
// ... or an ensemble holding one (the AsRef<GenericGraph> blanket impl of the crate)
template <class Ens>
auto as_graph(const Ens &e) -> decltype((e.graph)) { return e.graph; }

template <class G>
std::vector<double> vertex_load_cuda(Context &ctx, const G &graph_or_ensemble, bool include_endpoints) {
    DeviceGraph d(ctx, as_graph(graph_or_ensemble));
    return d.vertex_load(include_endpoints);
}

}  // namespace net_ensembles_cuda
