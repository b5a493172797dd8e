// Compiles the C++ mirror of the reference's operator interface (include/net_ensembles_cuda.hpp) against
// the C-ABI library and, when a GPU is present, runs the reference's own known-answer test for vertex_load
// (/root/reference tests/integration_test_graph.rs:238-275) through it.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/net_ensembles_cuda.hpp"

using namespace net_ensembles_cuda;

static bool close_enough(const std::vector<double> &a, const std::vector<double> &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::fabs(a[i] - b[i]) > 1e-9 * std::fmax(1.0, std::fabs(b[i]))) return false;
    return true;
}

int main(int argc, char **argv) {
    const bool compile_only = argc > 1 && !std::strcmp(argv[1], "--no-gpu");
    nec_host::Graph k4(4);
    for (uint32_t i = 0; i < 4; ++i)
        for (uint32_t j = i + 1; j < 4; ++j) k4.add_edge(i, j);
    nec_host::Graph g8(8);
    const uint32_t e[14][2] = {{0, 1}, {0, 2}, {2, 1}, {5, 1}, {2, 3}, {2, 4}, {4, 3}, {5, 3}, {4, 5}, {4, 7}, {4, 6}, {5, 7}, {5, 6}, {6, 7}};
    for (auto &p : e) g8.add_edge(p[0], p[1]);
    nec_host::ErEnsembleC erc(50, 4.0, nec_host::Pcg64::seed_from_u64(123239010));   // benches/bench_vertex_load.rs:6-9
    if (compile_only) {
        try {
            Context ctx(0);
        } catch (const std::exception &ex) {
            std::printf("no GPU: %s\n", ex.what());
            return std::strstr(ex.what(), "no CPU fallback") ? 0 : 2;
        }
        return 0;
    }
    Context ctx(0);
    if (!close_enough(vertex_load_cuda(ctx, k4, true), {3, 3, 3, 3})) return 10;
    if (!close_enough(vertex_load_cuda(ctx, k4, false), {0, 0, 0, 0})) return 11;
    if (!close_enough(vertex_load_cuda(ctx, nec_host::Graph(4), true), {0, 0, 0, 0})) return 12;
    if (!close_enough(vertex_load_cuda(ctx, g8, false),
                      {0.0, 4.666666666666666, 8.0, 0.6666666666666665, 8.666666666666668, 10.0, 0.0, 0.0})) return 13;
    {   // diameter known answers (tests/integration_test_graph.rs:166-184)
        nec_host::Graph path5(5);
        for (uint32_t i = 0; i < 4; ++i) path5.add_edge(i, i + 1);
        DeviceGraph d(ctx, path5);
        if (d.diameter().value_or(99) != 4) return 15;
        DeviceGraph k(ctx, k4);
        if (k.diameter().value_or(99) != 1) return 16;
        if (k.closeness_centrality() != std::vector<double>{1.0, 1.0, 1.0, 1.0}) return 17;
    }
    const std::vector<double> l = vertex_load_cuda(ctx, erc, true);   // ensemble -> graph via the blanket overload
    if (l.size() != 50) return 14;
    std::printf("cpp mirror ok\n");
    return 0;
}
