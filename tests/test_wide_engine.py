"""The compact batched engine (nec_wide.cuh; default for n > 180000) on graphs small enough for the oracle to see
every source: state re-use batch after batch, team sizes / CTA shapes / forward directions bit-identical, hand-over
of batches whose level queue does not fit, ragged batches and repeated sources.  The engine-parametrised tests of
test_gpu_parity.py run it on the golden vectors, fixtures and edge cases as well."""
import numpy as np
import pytest

import net_ensembles_b200 as ne
from tests.helpers import close

pytestmark = pytest.mark.gpu
REL = 1e-9


@pytest.fixture
def wide_ctx(ctx, monkeypatch):
    ctx.set_engine(3)
    monkeypatch.setenv("NEC_WIDE_QUEUE", "256")
    yield ctx
    ctx.set_workspace_limit(0)
    ctx.set_engine(0)


@pytest.mark.parametrize("width", ["64", "128", "256"])
def test_team_sizes_cta_shapes_and_directions_are_bitwise_identical(wide_ctx, oracle, monkeypatch, width):
    ctx = wide_ctx
    monkeypatch.setenv("NEC_WIDE_WIDTH", width)
    graphs = [ne.gnm_scalable(30000, 120000, 3).export_csr(), ne.ba_scalable(20000, 4, 5, 5).export_csr(),
              ne.ErEnsembleC(2500, 1.2, 8).export_csr(), ne.SwEnsemble(3000, 0.1, 5).export_csr()]
    for rp, col in graphs:
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        src = np.arange(0, n, max(1, n // 700), dtype=np.uint32)        # a few batches, the last one ragged
        outs = []
        for team, threads, pull, soft in (("1", "1024", "1", "1"), ("2", "1024", "1", "1"), ("3", "512", "1", "0"), ("8", "512", "0", "1"),
                                          ("4", "1024", "2", "0"), ("5", "1024", "1", "1"), ("2", "1024", "1", "0")):
            monkeypatch.setenv("NEC_WIDE_SOFT", soft)       # software teams (cooperative launch) or thread-block clusters
            monkeypatch.setenv("NEC_WIDE_TEAM", team)
            monkeypatch.setenv("NEC_WIDE_THREADS", threads)
            monkeypatch.setenv("NEC_PULL", pull)
            outs.append(dg.vertex_load_sources(src, False))
            st = ctx.stats()
            assert st["engine"] == 4 and st["batch_width"] == int(width) and st["cta_threads"] == int(threads)
            assert st["team_size"] <= int(team) and st["n_slots"] == st["n_batches"]      # one slot per batch: the slot sums are per batch
            if pull == "0":
                assert st["pull_levels"] == 0
            if pull == "2":
                assert st["push_levels"] == 0
        assert all(o.tobytes() == outs[0].tobytes() for o in outs)
        want = oracle.vertex_load(rp, col, False, sources=src, threads=0)
        assert close(outs[0], want, REL)
        assert st["reached_pairs"] == oracle.last_stats["reached_pairs"] and st["edge_pairs"] == oracle.last_stats["edge_pairs"]
        dg.free()


@pytest.mark.parametrize("width", ["64", "128", "256"])
def test_slots_are_reused_batch_after_batch(wide_ctx, oracle, monkeypatch, width):
    """Few slots, many batches per slot: records, level records, queues and runs of one batch must not leak into
    the next (only the tags are reset between batches)."""
    ctx = wide_ctx
    monkeypatch.setenv("NEC_WIDE_WIDTH", width)
    for e in (ne.ErEnsembleC(1800, 3.0, 77), ne.ba_scalable(5000, 3, 4, 9), ne.SwEnsemble(1500, 0.05, 3)):
        rp, col = e.export_csr()
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        free = dg.vertex_load(True)
        slots_free = ctx.stats()["n_slots"]
        ctx.set_workspace_limit(3 * (n * int(width) * 24 + (1 << 21)))          # room for a few slots
        few = dg.vertex_load(True)
        st = ctx.stats()
        assert st["engine"] == 4 and st["n_slots"] < slots_free and st["n_batches"] >= 3 * st["n_slots"]
        again = dg.vertex_load(True)
        assert again.tobytes() == few.tobytes()
        assert close(few, free, 1e-12) and close(few, oracle.vertex_load(rp, col, True, threads=0), REL)
        ctx.set_workspace_limit(0)
        dg.free()


def test_batches_that_do_not_fit_are_handed_to_the_source_minor_engine(wide_ctx, oracle, monkeypatch):
    """A slot has room for a dozen queue entries per vertex; a batch whose vertices sit on more distinct levels
    (long paths, lattices) overflows it, is reported, and re-run by the other engine: nothing is approximated."""
    ctx = wide_ctx
    ring = ne.SwEnsemble(1100, 0.0, 1)                       # un-rewired ring: a vertex sees 128 sources at ~128 distances
    rp, col = ring.export_csr()
    dg = ctx.upload(rp, col)
    monkeypatch.setenv("NEC_WIDE_QUEUE", "2")
    got = dg.vertex_load(True)
    st = ctx.stats()
    assert st["engine"] == (4 | 1 << 4) and st["requeued_batches"] > 0
    assert close(got, oracle.vertex_load(rp, col, True), REL)
    assert st["reached_pairs"] >= 1100 * 1100                # (the handed-over batches are counted by both engines' forward passes or only once)
    monkeypatch.setenv("NEC_WIDE_QUEUE", "256")
    full = dg.vertex_load(True)
    assert ctx.stats()["engine"] == 4 and ctx.stats()["requeued_batches"] == 0 and ctx.stats()["max_depth"] == 275
    assert close(full, got, REL)
    dg.free()


def test_ragged_batches_repeated_sources_and_components(wide_ctx, oracle):
    ctx = wide_ctx
    g = ne.ErEnsembleC(3000, 3.0, 11)                        # many components, isolated vertices
    rp, col = g.export_csr()
    dg = ctx.upload(rp, col)
    rng = np.random.default_rng(5)
    for count in (1, 37, 128, 129, 257, 128 * 5 + 37):
        src = rng.integers(0, 3000, count).astype(np.uint32)   # with repeats
        for incl in (True, False):
            got = dg.vertex_load_sources(src, incl)
            assert ctx.stats()["engine"] == 4
            assert close(got, oracle.vertex_load(rp, col, incl, sources=src, threads=0), REL), (count, incl)
    assert dg.vertex_load_sources(np.zeros(0, np.uint32), True).tolist() == [0.0] * 3000
    dg.free()


def test_differential_fuzz_against_the_oracle(wide_ctx, oracle, monkeypatch):
    rng = np.random.default_rng(2026)
    for trial in range(40):
        n = int(rng.integers(2, 400))
        m = int(rng.integers(0, 3 * n))
        edges = set()
        for _ in range(m):
            a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
            if a != b:
                edges.add((min(a, b), max(a, b)))
        adj = [[] for _ in range(n)]
        for a, b in sorted(edges, key=lambda e: rng.random()):
            adj[a].append(b)
            adj[b].append(a)
        rp = np.zeros(n + 1, np.uint64)
        rp[1:] = np.cumsum([len(a) for a in adj])
        col = np.array([x for a in adj for x in a], dtype=np.uint32)
        monkeypatch.setenv("NEC_WIDE_WIDTH", ("64", "128", "256")[trial % 3])
        monkeypatch.setenv("NEC_PULL", str((trial // 3) % 3))
        dg = wide_ctx.upload(rp, col)
        incl = bool(trial & 4)
        assert close(dg.vertex_load(incl), oracle.vertex_load(rp, col, incl), REL), trial
        dg.free()


def _csr_from_edges(n, a, b):
    """Undirected edge arrays -> CSR with rows in edge order (numpy only: these graphs have a few 10^5 vertices)."""
    src = np.concatenate([a, b]).astype(np.int64)
    dst = np.concatenate([b, a]).astype(np.uint32)
    order = np.argsort(src, kind="stable")
    rp = np.zeros(n + 1, np.uint64)
    rp[1:] = np.cumsum(np.bincount(src, minlength=n))
    return rp, dst[order]


def test_default_plan_on_a_giant_hub_with_a_tail_and_isolated_vertices(ctx, oracle):
    """Beyond the mid engine's size (default plan, no knobs): one hub with 200 000 leaves (a row of 200 000 entries: the pull
    scan's whole-warp gather, a backward tile of 200 000 rows), a path of 1500 vertices hanging off a leaf (1500 levels:
    32-bit level tags, one team barrier set per level), a second small component and isolated vertices."""
    hub, leaves, path, extra = 0, 200_000, 1500, 300
    n = 1 + leaves + path + extra + 50                       # the last 50 vertices are isolated
    a = [np.zeros(leaves, np.int64)]
    b = [np.arange(1, leaves + 1, dtype=np.int64)]
    p0 = 1 + leaves
    a.append(np.concatenate([[leaves], np.arange(p0, p0 + path - 1)]))          # last leaf -> path
    b.append(np.arange(p0, p0 + path))
    e0 = p0 + path
    a.append(np.arange(e0, e0 + extra - 1)); b.append(np.arange(e0 + 1, e0 + extra))   # second component: a short path ...
    a.append(np.arange(e0, e0 + extra - 2)); b.append(np.arange(e0 + 2, e0 + extra))   # ... with chords
    rp, col = _csr_from_edges(n, np.concatenate(a), np.concatenate(b))
    dg = ctx.upload(rp, col)
    src = np.concatenate([[hub, 1, 2, leaves, p0, p0 + path - 1, e0, e0 + extra - 1, n - 1],
                          np.arange(7, n, 9973)]).astype(np.uint32)
    for incl in (True, False):
        got = dg.vertex_load_sources(src, incl)
        st = ctx.stats()
        assert st["engine"] & 15 == 4 and st["max_depth"] >= path
        want = oracle.vertex_load(rp, col, incl, sources=src, threads=0)
        assert close(got, want, REL), incl
        assert st["reached_pairs"] == oracle.last_stats["reached_pairs"] and st["edge_pairs"] == oracle.last_stats["edge_pairs"]
    assert got[n - 1] == 0.0 and got[hub] > 1e5
    dg.free()


def test_default_plan_hands_a_lattice_back_and_the_deep_path_finishes_it(ctx, oracle):
    """A 450 x 450 grid through the default plan: a vertex sits on far more than a dozen distinct levels within a batch, so the
    batch overflows its slot's queue, is handed to the source-minor engine, and - 898 levels deep - to that engine's
    32-bit distance rows.  Nothing is approximated on the way."""
    side = 450
    n = side * side
    idx = np.arange(n, dtype=np.int64).reshape(side, side)
    a = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()])
    b = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()])
    rp, col = _csr_from_edges(n, a, b)
    dg = ctx.upload(rp, col)
    src = np.concatenate([[0, side - 1, n - 1, n // 2 + side // 2], np.arange(11, n, 3001)]).astype(np.uint32)
    got = dg.vertex_load_sources(src, True)
    st = ctx.stats()
    assert st["engine"] & 15 == 4 and st["requeued_batches"] >= 1 and st["max_depth"] == 2 * (side - 1)
    assert close(got, oracle.vertex_load(rp, col, True, sources=src, threads=0), REL)
    dg.free()


def test_planner_queue_room_follows_the_memory_it_is_given(ctx, oracle):
    """Room for 24 queue entries per vertex where it costs no speed, a dozen where the workspace is short - and either way
    the same loads (per-batch values do not depend on the queue's size), equal to the oracle's."""
    g = ne.gnm_scalable(30000, 120000, 9)
    rp, col = g.export_csr()
    n = len(rp) - 1
    dg = ctx.upload(rp, col)
    ctx.set_engine(3)
    src = np.arange(0, n, 5, dtype=np.uint32)                      # 6000 sources = 24 batches of 256
    roomy = dg.vertex_load_sources(src, True)
    st_roomy = ctx.stats()
    per_slot_roomy = st_roomy["workspace_bytes"] / st_roomy["n_slots"]
    limit = int(st_roomy["workspace_bytes"] * 0.3)                 # 7 slots with the long queue, 8 with the short one: one wave fewer
    ctx.set_workspace_limit(limit)
    tight = dg.vertex_load_sources(src, True)
    st_tight = ctx.stats()
    assert st_tight["engine"] == 4 and st_tight["requeued_batches"] == 0 and st_tight["n_slots"] < st_roomy["n_slots"]
    assert st_tight["workspace_bytes"] <= limit
    assert st_tight["workspace_bytes"] / st_tight["n_slots"] < per_slot_roomy          # the dozen
    assert close(tight, roomy, 1e-12) and close(roomy, oracle.vertex_load(rp, col, True, sources=src, threads=0), REL)
    ctx.set_workspace_limit(0)
    ctx.set_engine(0)
    dg.free()
