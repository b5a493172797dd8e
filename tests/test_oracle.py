"""Pins the CPU oracle (oracle/ref_vertex_load.c) to the reference's own known answers
(/root/reference tests/integration_test_graph.rs:238-275) and to independent cross-checks."""
import os
import sys

import numpy as np
import pytest

from tests.helpers import adj_from_edges, close, csr_from_adj, edge_case_graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from py_oracle import vertex_load as py_vertex_load  # noqa: E402


def test_reference_known_answers_bit_exact(oracle, golden):
    for ka in golden["known_answers"]:
        adj = adj_from_edges(ka["n"], ka["edges_in_insertion_order"])
        rp, col = csr_from_adj(adj)
        for key, incl in (("include_endpoints_true", True), ("include_endpoints_false", False)):
            if key in ka:
                got = oracle.vertex_load(rp, col, incl)
                assert got.tolist() == ka[key], (ka["name"], key)      # exact, like assert_eq! in the crate


def test_c_oracle_equals_python_restatement(oracle):
    for name, adj in edge_case_graphs().items():
        rp, col = csr_from_adj(adj)
        for incl in (True, False):
            assert oracle.vertex_load(rp, col, incl).tolist() == py_vertex_load(adj, incl), name


def test_fixture_loads_match_committed_derivation(oracle, golden):
    for name, g in golden["graphs"].items():
        rp, col = csr_from_adj(g["adj"])
        d = golden["derived_load"][name]
        t, f = oracle.vertex_load(rp, col, True), oracle.vertex_load(rp, col, False)
        assert [float(x).hex() for x in t] == d["true_hex"], name
        assert [float(x).hex() for x in f] == d["false_hex"], name
        # sum of load_true = sum of BFS distances over ordered reachable pairs: an integer
        assert t.sum() == pytest.approx(d["sum_true"], rel=1e-13) and abs(d["sum_true"] - round(d["sum_true"])) < 1e-6
        assert int(t.argmax()) == d["argmax_true"]


def test_oracle_vs_networkx_load_centrality(oracle, golden):
    nx = pytest.importorskip("networkx")
    for name, g in golden["graphs"].items():
        G = nx.Graph()
        G.add_nodes_from(range(g["n"]))
        G.add_edges_from((i, j) for i, a in enumerate(g["adj"]) for j in a)
        lc = nx.load_centrality(G, normalized=False)
        rp, col = csr_from_adj(g["adj"])
        ours = oracle.vertex_load(rp, col, False)
        ref = np.array([lc[i] for i in range(g["n"])])
        assert close(ours, ref, rel=1e-12, abs_floor=1e-9), name


def test_load_is_not_brandes_betweenness(oracle):
    """The crate splits equally among predecessors (generic_graph.rs:1377); sigma must not leak in."""
    adj = edge_case_graphs()["load_vs_brandes7"]
    rp, col = csr_from_adj(adj)
    got = oracle.vertex_load(rp, col, False)
    expect = [8.1667, 2.1667, 2.1667, 4.8333, 8.1667, 3.6667, 4.8333]
    assert np.allclose(got, expect, atol=1e-4)
    brandes = [8.3333, 2.3333, 2.3333, 4.6667, 8.3333, 3.3333, 4.6667]
    assert not np.allclose(got, brandes, atol=1e-2)


def test_invariants(oracle, golden):
    for name, g in golden["graphs"].items():
        rp, col = csr_from_adj(g["adj"])
        t, f = oracle.vertex_load(rp, col, True), oracle.vertex_load(rp, col, False)
        # load_true - load_false = |component| - 1
        comp = np.array([np.count_nonzero(oracle.bfs_debug(rp, col, v)[0] != 0xFFFFFFFF) for v in range(g["n"])])
        assert close(t - f, comp - 1, rel=1e-12), name
        # sum_v load_true(v) = sum over ordered reachable pairs of d(s,t)
        tot = 0
        for s in range(g["n"]):
            d = oracle.bfs_debug(rp, col, s)[0]
            tot += int(d[d != 0xFFFFFFFF].sum())
        assert abs(t.sum() - tot) <= 1e-9 * tot, name


def test_sources_subset_and_threads(oracle, golden):
    g = golden["graphs"]["erc_2"]
    rp, col = csr_from_adj(g["adj"])
    full = oracle.vertex_load(rp, col, True)
    a = oracle.vertex_load(rp, col, True, sources=np.arange(0, g["n"], 2))
    b = oracle.vertex_load(rp, col, True, sources=np.arange(1, g["n"], 2))
    assert close(a + b, full, rel=1e-12)
    mt = oracle.vertex_load(rp, col, True, threads=3)
    assert close(mt, full, rel=1e-12)
    assert oracle.last_stats["reached_pairs"] > 0


def test_bfs_debug_small(oracle):
    adj = edge_case_graphs()["load_vs_brandes7"]
    rp, col = csr_from_adj(adj)
    dist, npred, sigma, bk = oracle.bfs_debug(rp, col, 0)
    assert dist.tolist() == [0, 1, 1, 1, 2, 2, 3]
    assert npred.tolist() == [0, 1, 1, 1, 2, 1, 2]
    assert sigma.tolist() == [1.0, 1.0, 1.0, 1.0, 2.0, 1.0, 3.0]
