"""The crate's serde_json save-states as an input format of the path: topology in storage order,
parameters and generator state.  CPU tests: round trips through this package's writer, the reduced golden fixtures, and -
in the build container, where /root/reference exists - the reference's own TestData dumps.  One GPU test: a loaded
topology gives bitwise the load of the ensemble it was saved from (same adjacency order, same CSR)."""
import json
import os

import numpy as np
import pytest

import net_ensembles_b200 as ne

MAKERS = {"erc": lambda g: ne.ErEnsembleC(int(g["n"]), float(g["c"]), int(g["seed"])),
          "erm": lambda g: ne.ErEnsembleM(int(g["n"]), int(g["m"]), int(g["seed"])),
          "sw": lambda g: ne.SwEnsemble(int(g["n"]), float(g["r_prob"]), int(g["seed"]))}


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ["erc_1", "erc_2", "erm_1", "erm_2", "sw_1", "sw_2", "sw_3", "sw_4"])
def test_dump_and_load_round_trip_equals_the_golden_fixture(built, golden, name, tmp_path):
    g = golden["graphs"][name]
    e = MAKERS[g["kind"]](g)
    path = tmp_path / (name + ".json")
    doc = ne.dump_serde_json(e, str(path))
    assert doc["graph"]["next_id"] == int(g["n"]) and doc["graph"]["edge_count"] == int(g["edge_count"]) and doc["graph"]["phantom"] is None
    for src in (str(path), doc, json.dumps(doc), open(path)):
        st = ne.load_serde_json(src)
        assert st.kind == g["kind"]
        assert st.graph.adjacency() == g["adj"] and st.graph.edge_count() == int(g["edge_count"])
        assert st.rng == (int(g["rng_state"]), int(g["rng_increment"]))
    if g["kind"] == "erc":
        assert st.params == {"prob": float(g["c"]) / (int(g["n"]) - 1), "c_target": float(g["c"])}
    if g["kind"] == "sw":
        flat = [t for row in g["root_target"] for t in row]
        assert st.root_targets.tolist() == flat and st.params == {"r_prob": float(g["r_prob"])}
    else:
        assert st.root_targets is None


@pytest.mark.skipif(not os.path.isdir("/root/reference/TestData"), reason="the reference checkout exists in the build container only")
@pytest.mark.parametrize("name", ["erc_1", "erc_2", "erm_1", "erm_2", "sw_1", "sw_2", "sw_3", "sw_4"])
def test_reference_dumps_load_and_equal_the_golden_fixture(built, golden, name):
    g = golden["graphs"][name]
    st = ne.load_serde_json("/root/reference/TestData/unchaning_%s.json" % name)
    assert st.kind == g["kind"] and st.graph.adjacency() == g["adj"]
    assert st.rng == (int(g["rng_state"]), int(g["rng_increment"]))
    if g["kind"] == "sw":
        assert st.root_targets.tolist() == [t for row in g["root_target"] for t in row]
    if g["kind"] == "erm":
        assert st.params == {"m": int(g["m"])}


def test_bare_graph_and_malformed_dumps(built):
    g = ne.Graph(5)
    for a, b in ((0, 3), (3, 1), (1, 0), (4, 3)):
        g.add_edge(a, b)
    doc = ne.dump_serde_json(g)
    assert "graph" not in doc and [v["adj"] for v in doc["vertices"]] == [[3, 1], [3, 0], [], [0, 1, 4], [3]]
    st = ne.load_serde_json(doc)
    assert st.kind == "graph" and st.rng is None and st.graph.adjacency() == g.adjacency()
    bad = json.loads(json.dumps(doc))
    bad["vertices"][0]["adj"] = [3]                              # (1, 0) has lost its counterpart
    with pytest.raises(ValueError, match="counterpart"):
        ne.load_serde_json(bad)
    bad = json.loads(json.dumps(doc))
    bad["vertices"][2]["adj"] = [2]
    with pytest.raises(ValueError, match="self-loop"):
        ne.load_serde_json(bad)
    bad = json.loads(json.dumps(doc))
    bad["vertices"][0]["adj"] = [3, 1, 7]
    with pytest.raises(ValueError, match="out of range"):
        ne.load_serde_json(bad)
    bad = json.loads(json.dumps(doc))
    bad["edge_count"] = 9
    with pytest.raises(ValueError, match="edge_count"):
        ne.load_serde_json(bad)
    bad = json.loads(json.dumps(doc))
    bad["vertices"][1]["id"] = 4
    with pytest.raises(ValueError, match="carries id"):
        ne.load_serde_json(bad)
    with pytest.raises(ValueError, match="vertices"):
        ne.load_serde_json({"rng": {}})
    assert ne.graph_from_adjacency([]).vertex_count() == 0


@pytest.mark.gpu
def test_loaded_topology_gives_bitwise_the_load_of_its_ensemble(ctx, oracle, tmp_path):
    for e in (ne.ErEnsembleC(500, 3.0, 11), ne.SwEnsemble(700, 0.2, 5)):
        e.m_steps(50)
        st = ne.load_serde_json(ne.dump_serde_json(e, str(tmp_path / "state.json")))
        mine, theirs = st.vertex_load(True, ctx), e.vertex_load(True, ctx)
        assert mine.tobytes() == theirs.tobytes()
        rp, col = e.export_csr()
        want = oracle.vertex_load(rp, col, True)
        assert float(np.max(np.abs(mine - want) / np.maximum(np.abs(want), 1.0))) <= 1e-9


@pytest.mark.parametrize("name", ["erc_1", "erm_1", "sw_1", "sw_2", "sw_4"])
def test_resumed_ensemble_continues_like_the_one_it_was_saved_from(built, golden, name, tmp_path):
    """dump -> load -> resume, then both copies take the same Markov steps and the same randomize(): lists (order included),
    small-world root targets, ErM's edge lists and the generator stay identical."""
    g = golden["graphs"][name]
    original = MAKERS[g["kind"]](g)
    original.m_steps(37)
    copy = ne.load_serde_json(ne.dump_serde_json(original, str(tmp_path / "s.json"))).resume()
    assert type(copy) is type(original) and copy.rng_state() == original.rng_state() and copy.adjacency() == original.adjacency()
    for action in (lambda e: e.m_steps(500), lambda e: e.randomize(), lambda e: e.m_steps(123)):
        action(original), action(copy)
        assert copy.adjacency() == original.adjacency() and copy.rng_state() == original.rng_state()
        if g["kind"] == "sw":
            assert copy.root_targets().tolist() == original.root_targets().tolist()
    assert ne.dump_serde_json(copy) == ne.dump_serde_json(original)


@pytest.mark.skipif(not os.path.isdir("/root/reference/TestData"), reason="the reference checkout exists in the build container only")
@pytest.mark.parametrize("name", ["erc_1", "erc_2", "erm_1", "erm_2", "sw_1", "sw_2", "sw_3", "sw_4"])
def test_reference_dumps_resume_into_the_state_the_seeded_constructor_gives(built, golden, name):
    """The reference's TestData dumps are the state right after construction: resuming one must give an ensemble that
    is indistinguishable from ours built from the same seed - including ErM's private edge lists - now and after steps."""
    g = golden["graphs"][name]
    with open("/root/reference/TestData/unchaning_%s.json" % name) as fh:
        raw = json.load(fh)
    resumed, built_here = ne.load_serde_json(raw).resume(), MAKERS[g["kind"]](g)
    mine = ne.dump_serde_json(built_here)
    for key in ("all_edges", "possible_edges", "current_edges", "rng", "m", "r_prob", "c_target", "prob"):
        assert mine.get(key) == raw.get(key), key
    assert [v["adj"] for v in mine["graph"]["vertices"]] == [v["adj"] for v in raw["graph"]["vertices"]]
    for e in (resumed, built_here):
        e.m_steps(300)
    assert resumed.adjacency() == built_here.adjacency() and resumed.rng_state() == built_here.rng_state()


def test_resume_rejects_states_that_do_not_fit(built):
    e = ne.SwEnsemble(30, 0.3, 3)
    doc = ne.dump_serde_json(e)
    broken = json.loads(json.dumps(doc))
    for entry in broken["graph"]["vertices"][0]["adj"]:
        entry["originally_to"] = None                              # vertex 0 loses its root edges
    with pytest.raises(ValueError, match="root"):
        ne.load_serde_json(broken).resume()
    m = ne.dump_serde_json(ne.ErEnsembleM(12, 20, 5))
    m["current_edges"] = m["current_edges"][:-1]
    with pytest.raises(ValueError, match="edge lists"):
        ne.load_serde_json(m).resume()
    with pytest.raises(ValueError, match="no ensemble"):
        ne.load_serde_json(ne.dump_serde_json(ne.Graph(4))).resume()


@pytest.mark.gpu
def test_resumed_chains_run_device_resident_like_their_originals(ctx, tmp_path):
    """Checkpoint / resume around the batched Markov-chain path: chains saved, resumed and put into a ChainBatch give
    bitwise the loads of the chains they were saved from, step after step."""
    originals = [ne.ErEnsembleC(96, 4.0, 500 + g) for g in range(40)]
    for e in originals:
        e.m_steps(10)
    copies = [ne.load_serde_json(ne.dump_serde_json(e)).resume() for e in originals]
    a, b = ne.ChainBatch(originals), ne.ChainBatch(copies)
    for _ in range(3):
        ra, rb = a.step_and_load(ctx, 5, True), b.step_and_load(ctx, 5, True)
        assert ra.tobytes() == rb.tobytes()
    assert [e.adjacency() for e in originals] == [e.adjacency() for e in copies]
    a.close(), b.close()
