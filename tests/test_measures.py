"""the all-source BFS measures beside vertex_load — diameter,
longest_shortest_path_from_index, closeness_centrality (/root/reference
src/generic_graph/generic_graph.rs:1116-1144, :1387-1403).  Integer work: parity is exact."""
import numpy as np
import pytest

import net_ensembles_b200 as ne
from tests.helpers import adj_from_edges, csr_from_adj, edge_case_graphs


def _reference_diameter_cases():
    # tests/integration_test_graph.rs:166-196
    return [
        (adj_from_edges(5, [(i, i + 1) for i in range(4)]), 4),
        (adj_from_edges(4, [(i, (i + 1) % 4) for i in range(4)]), 2),
        (adj_from_edges(40, [(i, (i + 1) % 40) for i in range(40)]), 20),
        (adj_from_edges(40, [(i, j) for i in range(40) for j in range(i + 1, 40)]), 1),
        ([[]], 0),
    ]


def test_oracle_diameter_known_answers(oracle):
    for adj, want in _reference_diameter_cases():
        assert oracle.diameter(*csr_from_adj(adj)) == want
    assert oracle.diameter(*csr_from_adj([[], []])) is None             # not connected
    assert oracle.diameter(np.zeros(1, np.uint64), np.zeros(0, np.uint32)) is None   # no vertices


def test_oracle_measures_are_consistent(oracle, golden):
    for name, g in golden["graphs"].items():
        rp, col = csr_from_adj(g["adj"])
        ecc, sumd, reached = oracle.distance_measures(rp, col)
        clo = oracle.closeness_centrality(rp, col)
        with np.errstate(divide="ignore"):
            assert np.array_equal(clo, np.float64(g["n"] - 1) / sumd.astype(np.float64)), name   # symmetry of d(s,t)
        d = oracle.diameter(rp, col)
        assert d == (int(ecc.max()) if reached[0] == g["n"] else None), name
        # sum of all distances = sum_v load_true(v) (the vertex_load invariant), cross-check with the other oracle
        assert int(sumd.sum()) == int(round(oracle.vertex_load(rp, col, True).sum())), name


@pytest.mark.gpu
def test_gpu_diameter_known_answers(ctx):
    for adj, want in _reference_diameter_cases():
        dg = ctx.upload(*csr_from_adj(adj))
        assert dg.diameter() == want
        dg.free()
    dg = ctx.upload(*csr_from_adj([[], []]))
    assert dg.diameter() is None
    dg.free()
    assert ctx.upload(np.zeros(1, np.uint64), np.zeros(0, np.uint32)).diameter() is None


@pytest.mark.gpu
def test_gpu_measures_exact_on_fixtures_and_edge_cases(ctx, oracle, golden):
    graphs = {k: g["adj"] for k, g in golden["graphs"].items()}
    graphs.update(edge_case_graphs())
    for name, adj in graphs.items():
        rp, col = csr_from_adj(adj)
        dg = ctx.upload(rp, col)
        ecc, sumd, reached = dg.distance_measures()
        recc, rsum, rreach = oracle.distance_measures(rp, col)
        assert ecc.tolist() == recc.tolist() and sumd.tolist() == rsum.tolist() and reached.tolist() == rreach.tolist(), name
        assert dg.diameter() == oracle.diameter(rp, col), name
        got, want = dg.closeness_centrality(), oracle.closeness_centrality(rp, col)
        assert np.array_equal(got, want, equal_nan=True), name             # same integer sums, same f64 division
        assert dg.longest_shortest_path_from_index(len(adj) - 1) == int(recc[-1]), name
        dg.free()


@pytest.mark.gpu
def test_gpu_measures_on_ensembles_and_deep_graphs(ctx, oracle):
    for e in (ne.SwEnsemble(3000, 0.1, 5), ne.ErEnsembleC(2000, 3.0, 8), ne.SwEnsemble(1100, 0.0, 1)):   # last: depth 275
        rp, col = e.export_csr()
        assert e.diameter(ctx) == oracle.diameter(rp, col)
        assert np.array_equal(e.closeness_centrality(ctx), oracle.closeness_centrality(rp, col), equal_nan=True)
    ring = ne.SwEnsemble(4096, 0.0, 1)           # mid engine hands every source back (depth 1024 > 253)
    rp, col = ring.export_csr()
    dg = ctx.upload(rp, col)
    got, want = dg.distance_measures(), oracle.distance_measures(rp, col)
    assert all(a.tolist() == b.tolist() for a, b in zip(got, want))
    assert dg.diameter() == 1024
    dg.free()
    g = ne.gnm_scalable(1 << 15, 1 << 17, 3)
    rp, col = g.export_csr()
    dg = ctx.upload(rp, col)
    src = np.arange(0, 1 << 15, 331, dtype=np.uint32)
    got, want = dg.distance_measures(src), oracle.distance_measures(rp, col, src)
    assert all(a.tolist() == b.tolist() for a, b in zip(got, want))
    dg.free()
