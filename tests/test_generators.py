"""The host-side ensemble mirror against the reference's seeded fixtures
(/root/reference TestData/unchaning_*.json, reduced into tests/golden/fixtures.json):
adjacency ORDER and the final Pcg64 state must be bit-identical."""
import numpy as np
import pytest

import net_ensembles_b200 as ne


def _build(g):
    if g["kind"] == "erc":
        return ne.ErEnsembleC(g["n"], g["c"], g["seed"])
    if g["kind"] == "erm":
        return ne.ErEnsembleM(g["n"], g["m"], g["seed"])
    return ne.SwEnsemble(g["n"], g["r_prob"], g["seed"])


@pytest.mark.parametrize("name", ["erc_1", "erc_2", "erm_1", "erm_2", "sw_1", "sw_2", "sw_3", "sw_4"])
def test_seeded_construction_matches_reference_fixture(built, golden, name):
    g = golden["graphs"][name]
    e = _build(g)
    assert e.adjacency() == g["adj"]
    assert e.edge_count() == g["edge_count"]
    state, inc = e.rng_state()
    assert str(state) == g["rng_state"] and str(inc) == g["rng_increment"]
    if g["kind"] == "sw":
        flat = [x if x >= 0 else 0xFFFFFFFF for row in g["root_target"] for x in row]
        assert e.root_targets().tolist() == flat


def test_erc_edge_counts_of_reference_test(built):
    # tests/integration_test_er_c.rs:98-112: ErEnsembleC(20, 2.7, seed 76) has 28 edges
    e = ne.ErEnsembleC(20, 2.7, 76)
    assert e.edge_count() == 28


def test_graph_edge_semantics(built):
    g = ne.Graph(5)
    assert g.add_edge(0, 1) and g.add_edge(0, 2) and g.add_edge(3, 0)
    assert not g.add_edge(1, 0)                       # GraphErrors::EdgeExists
    assert g.adjacency() == [[1, 2, 3], [0], [0], [0], []]
    assert g.remove_edge(0, 1)                        # swap_remove: last element moves into the hole
    assert g.adjacency() == [[3, 2], [], [0], [0], []]
    assert not g.remove_edge(0, 1)                    # GraphErrors::EdgeDoesNotExist
    assert g.edge_count() == 2
    with pytest.raises(ValueError):
        g.add_edge(0, 9)


def _is_simple_undirected(adj):
    s = [set(a) for a in adj]
    return all(len(s[i]) == len(adj[i]) and i not in s[i] and all(i in s[j] for j in s[i]) for i in range(len(adj)))


def test_markov_steps_keep_graphs_valid(built):
    erm = ne.ErEnsembleM(40, 100, 5)
    sw = ne.SwEnsemble(60, 0.2, 5)
    erc = ne.ErEnsembleC(50, 4.0, 5)
    for _ in range(20):
        erm.m_steps(10), sw.m_steps(10), erc.m_steps(10)
        assert erm.edge_count() == 100 and _is_simple_undirected(erm.adjacency())
        assert sw.edge_count() == 120 and _is_simple_undirected(sw.adjacency())
        assert _is_simple_undirected(erc.adjacency())
    # every small-world vertex keeps exactly two root edges
    rp, _ = sw.export_csr()
    roots = sw.root_targets() != 0xFFFFFFFF
    assert all(roots[int(rp[i]):int(rp[i + 1])].sum() == 2 for i in range(60))


def test_ba_and_scalable_generators(built):
    ba = ne.BAensemble(300, 11, 3, 4)
    adj = ba.adjacency()
    assert _is_simple_undirected(adj) and min(len(a) for a in adj) >= 3
    assert ba.edge_count() == 4 * 3 // 2 + 3 * (300 - 4)
    g = ne.gnm_scalable(5000, 20000, 3)
    assert g.edge_count() == 20000 and _is_simple_undirected(g.adjacency())
    b = ne.ba_scalable(5000, 4, 5, 3)
    adj = b.adjacency()
    assert _is_simple_undirected(adj) and b.edge_count() == 10 + 4 * (5000 - 5)
    assert max(len(a) for a in adj) > 40              # heavy tail


def test_pcg64_known_answer(built):
    """Pcg64::seed_from_u64(123929) run through ErEnsembleC(123, 2.0) ends in the fixture's state;
    additionally the raw stream is reproducible and distinct per seed."""
    from net_ensembles_b200 import _lib
    lib = _lib.load_host()
    a, b = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    lib.nech_pcg64_draws(1, 4, a.ctypes.data)
    lib.nech_pcg64_draws(1, 4, b.ctypes.data)
    assert a.tolist() == b.tolist()
    lib.nech_pcg64_draws(2, 4, b.ctypes.data)
    assert a.tolist() != b.tolist()


def test_adjacency_entries_are_twice_the_edge_count_after_any_move():
    """The chain exporter sizes a CSR as 2 * edge_count without walking the adjacency lists (and checks it
    while exporting): the containers must keep the count through every kind of move."""
    import net_ensembles_b200 as ne
    ens = [ne.ErEnsembleC(200, 3.0, 1), ne.ErEnsembleM(150, 400, 2), ne.SwEnsemble(300, 0.3, 3), ne.BAensemble(120, 3, 4, 5)]
    for e in ens:
        for _ in range(3):
            rp, col = e.export_csr()
            assert len(col) == 2 * e.edge_count() and int(rp[-1]) == len(col)      # export_csr sizes by walking the lists
            e.randomize()
            if not isinstance(e, ne.BAensemble):
                e.m_steps(500)
    g = ne.Graph(10)
    assert g.add_edge(0, 1) and g.add_edge(1, 2) and not g.add_edge(0, 1) and g.remove_edge(0, 1) and not g.remove_edge(0, 1)
    assert len(g.export_csr()[1]) == 2 * g.edge_count() == 2
