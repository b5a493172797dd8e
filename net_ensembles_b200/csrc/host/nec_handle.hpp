// The handle the host-side C entry points pass around: one graph or ensemble of the mirror in
// nec_graphs.hpp.  Shared by libnec_host.so (generators, stepping, CSR export — no CUDA) and by the
// bridge compiled into the CUDA library (the drop-in vertex_load call over such a handle).
#pragma once
#include <memory>

#include "nec_graphs.hpp"

namespace nec_host {

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
    // one Markov step; false if this kind has no Markov chain (plain graphs, BA)
    bool m_step() {
        switch (kind) {
            case Kind::ErC: erc->m_step(); return true;
            case Kind::ErM: erm->m_step(); return true;
            case Kind::Sw: sw->m_step(); return true;
            default: return false;
        }
    }
};

template <class F>
auto with_graph(const Handle *h, F &&f) {
    if (h->kind == Kind::Sw) return f(h->sw->graph);
    return f(*h->plain_graph());
}

}  // namespace nec_host
