"""Parity tests proper: the CUDA path, called through the C ABI, against the pinned oracle on
identical inputs.  Integer state (distances, predecessor counts) must be exact; path counts are
exact while they fit in f64; load values agree within relative 1e-9 (absolute floor 1e-9 for
|x| < 1) — the tolerance BASELINE.json's north_star states, because the GPU summation order
(pull, lane tree, per-slot partials) differs from the crate's push order."""
import os

import numpy as np
import pytest

import net_ensembles_b200 as ne
from tests.helpers import adj_from_edges, close, csr_from_adj, edge_case_graphs

pytestmark = pytest.mark.gpu
REL = 1e-9
ENGINES = {"auto": 0, "batched": 1, "mid": 2, "wide": 3}


@pytest.fixture(params=["batched", "batched64", "batched_pull", "batched64_pull", "batched64_push", "mid",
                        "wide128", "wide64", "wide128_pull", "wide128_push", "wide64_pull", "wide128_t512", "wide64_team3",
                        "wide256", "wide256_pull", "wide256_push", "wide256_team3", "wide128_cluster", "wide256_cluster"])
def engine_ctx(ctx, request, monkeypatch):
    """Runs a test once per traversal engine, batch width and forward direction (the C ABI picks by graph size,
    source count and a per-level cost model by default; NEC_BATCH_WIDTH pins the source-minor batched engine's 32 / 64
    sources per batch, NEC_WIDE_WIDTH the compact engine's 64 / 128 / 256, NEC_PULL the forward pass to push only (0) or pull
    only (2); t512 = 512-thread CTAs, team3 = three CTAs per slot, cluster = thread-block clusters instead of the
    default software teams)."""
    name, _, variant = request.param.partition("_")
    base = name.rstrip("0123456789")
    ctx.set_engine(ENGINES[base])
    if base == "batched":
        monkeypatch.setenv("NEC_BATCH_WIDTH", "64" if name.endswith("64") else "32")
    if base == "wide":
        monkeypatch.setenv("NEC_WIDE_WIDTH", name[4:])
        monkeypatch.setenv("NEC_WIDE_QUEUE", "256")        # room for every level a vertex can sit on: nothing is handed to the other engine
    if variant in ("pull", "push"):
        monkeypatch.setenv("NEC_PULL", "2" if variant == "pull" else "0")
    if variant == "t512":
        monkeypatch.setenv("NEC_WIDE_THREADS", "512")
    if variant == "team3":
        monkeypatch.setenv("NEC_WIDE_TEAM", "3")
    if variant == "cluster":                               # thread-block clusters instead of software teams
        monkeypatch.setenv("NEC_WIDE_SOFT", "0")
    ctx.engine_name = base
    yield ctx
    ctx.set_engine(0)


def _gpu_load(ctx, adj, incl):
    rp, col = csr_from_adj(adj)
    g = ctx.upload(rp, col)
    try:
        return g.vertex_load(incl)
    finally:
        g.free()


def test_reference_known_answers(engine_ctx, golden):
    ctx = engine_ctx
    for ka in golden["known_answers"]:
        adj = adj_from_edges(ka["n"], ka["edges_in_insertion_order"])
        for key, incl in (("include_endpoints_true", True), ("include_endpoints_false", False)):
            if key in ka:
                got = _gpu_load(ctx, adj, incl)
                assert close(got, ka[key], REL), (ka["name"], key, got)
    # integers are exact whatever the summation order
    k4 = adj_from_edges(4, [(i, j) for i in range(4) for j in range(i + 1, 4)])
    assert _gpu_load(ctx, k4, True).tolist() == [3.0] * 4
    assert _gpu_load(ctx, k4, False).tolist() == [0.0] * 4


@pytest.mark.parametrize("name", ["erc_1", "erc_2", "erm_1", "erm_2", "sw_1", "sw_2", "sw_3", "sw_4"])
def test_reference_fixture_graphs(engine_ctx, oracle, golden, name):
    ctx = engine_ctx
    g = golden["graphs"][name]
    rp, col = csr_from_adj(g["adj"])
    dg = ctx.upload(rp, col)
    for incl in (True, False):
        got = dg.vertex_load(incl)
        want = np.array([float.fromhex(h) for h in golden["derived_load"][name]["true_hex" if incl else "false_hex"]])
        assert close(got, want, REL), name
        assert close(got, oracle.vertex_load(rp, col, incl), REL), name
    st = ctx.stats()
    assert st["n_sources"] == g["n"] and st["kernel_launches"] >= 2
    oracle.vertex_load(rp, col, True)
    assert st["edge_pairs"] == oracle.last_stats["edge_pairs"]
    assert st["reached_pairs"] == oracle.last_stats["reached_pairs"]
    dg.free()


@pytest.mark.parametrize("name", sorted(edge_case_graphs()))
def test_edge_cases(engine_ctx, oracle, name):
    ctx = engine_ctx
    adj = edge_case_graphs()[name]
    rp, col = csr_from_adj(adj)
    for incl in (True, False):
        assert close(_gpu_load(ctx, adj, incl), oracle.vertex_load(rp, col, incl), REL), name


def test_empty_graph(ctx):
    g = ctx.upload(np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    assert g.vertex_load(True).shape == (0,)


def test_ensemble_drop_in_call(engine_ctx, oracle):
    ctx = engine_ctx
    """ensemble.vertex_load(include_endpoints) — same call shape as the crate's trait method."""
    for e in (ne.ErEnsembleC(1000, 4.0, 123_239_010), ne.ErEnsembleM(200, 700, 4), ne.SwEnsemble(500, 0.1, 9),
              ne.BAensemble(400, 3, 2, 3), ne.ErEnsembleC(50, 4.0, 123_239_010)):
        rp, col = e.export_csr()
        for incl in (True, False):
            assert close(e.vertex_load(incl, ctx), oracle.vertex_load(rp, col, incl), REL)


def test_intermediate_state_exact(ctx, oracle, golden):
    for name in ("erc_1", "erm_1", "sw_1", "sw_3"):
        g = golden["graphs"][name]
        rp, col = csr_from_adj(g["adj"])
        dg = ctx.upload(rp, col)
        for s in (0, 7, g["n"] - 1):
            dist, npred, sigma, bk = dg.bfs_debug(s)
            rd, rn, rs, rb = oracle.bfs_debug(rp, col, s)
            assert dist.tolist() == rd.tolist(), (name, s)
            assert npred.tolist() == rn.tolist(), (name, s)
            assert sigma.tolist() == rs.tolist(), (name, s)     # exact: counts fit in f64 here
            assert close(bk, rb, REL), (name, s)
        dg.free()


def test_sources_subset_is_the_sharding_primitive(engine_ctx, oracle):
    ctx = engine_ctx
    e = ne.ErEnsembleC(700, 3.0, 17)
    rp, col = e.export_csr()
    dg = ctx.upload(rp, col)
    full = dg.vertex_load(True)
    parts = [dg.vertex_load_sources(np.arange(r, 700, 3), True) for r in range(3)]
    assert close(sum(parts), full, REL)
    sub = np.array([5, 99, 100, 640, 5], dtype=np.uint32)          # duplicates count twice
    assert close(dg.vertex_load_sources(sub, False), oracle.vertex_load(rp, col, False, sources=sub), REL)
    assert dg.vertex_load_sources(np.zeros(0, np.uint32), True).tolist() == [0.0] * 700
    with pytest.raises(ne.NecError):
        dg.vertex_load_sources(np.array([700], np.uint32), True)
    dg.free()


def test_deep_graphs_take_the_wide_distance_path(engine_ctx, oracle):
    ctx = engine_ctx
    """BFS depth >= 254 does not fit the 8-bit distance rows; those batches re-run with u32."""
    n = 700
    path = adj_from_edges(n, [(i, i + 1) for i in range(n - 1)])
    rp, col = csr_from_adj(path)
    dg = ctx.upload(rp, col)
    got = dg.vertex_load(True)
    assert (ctx.stats()["wide_batches"] > 0 or ctx.engine_name == "wide") and ctx.stats()["max_depth"] == n - 1
    assert close(got, oracle.vertex_load(rp, col, True), REL)
    # path: a source left of v sends the n-v vertices at/behind v through it, one right of v sends v+1
    v = np.arange(n)
    assert close(got, 1.0 * (v * (n - v) + (n - 1 - v) * (v + 1)), REL)
    dist = dg.bfs_debug(0)[0]
    assert dist.tolist() == list(range(n))
    dg.free()
    ring = ne.SwEnsemble(1100, 0.0, 1)                       # un-rewired ring-2: depth n/4 = 275
    rp, col = ring.export_csr()
    assert close(ring.vertex_load(False, ctx), oracle.vertex_load(rp, col, False), REL)


def test_run_to_run_bitwise_reproducible(engine_ctx):
    ctx = engine_ctx
    e = ne.SwEnsemble(3000, 0.1, 5)
    dg = e.to_device(ctx)
    a, b = dg.vertex_load(True), dg.vertex_load(True)
    assert a.tobytes() == b.tobytes()
    dg.free()


def test_batch_of_small_graphs_equals_singles(ctx, oracle):
    """One CTA per graph, one warp per source: equal to nec_vertex_load on each graph up to the order in
    which sources are summed (<= 1e-12), bit-reproducible run to run."""
    ens = [ne.ErEnsembleC(256, 4.0, 123_239_010 + g) for g in range(24)]
    ens += [ne.ErEnsembleC(33, 2.0, 3), ne.SwEnsemble(100, 0.2, 8), ne.ErEnsembleM(64, 200, 2), ne.ErEnsembleC(512, 3.0, 77)]
    graphs = [e.export_csr() for e in ens]
    graphs.append(csr_from_adj([[]]))                                   # n = 1
    graphs.append(csr_from_adj(adj_from_edges(300, [(i, i + 1) for i in range(299)])))  # depth 299
    for incl in (True, False):
        outs = ne.vertex_load_batch(ctx, [(rp.astype(np.uint32), col) for rp, col in graphs], incl)
        again = ne.vertex_load_batch(ctx, [(rp.astype(np.uint32), col) for rp, col in graphs], incl)
        for (rp, col), got, got2 in zip(graphs, outs, again):
            dg = ctx.upload(rp, col)
            single = dg.vertex_load(incl)
            dg.free()
            assert got.tobytes() == got2.tobytes()
            assert close(got, single, 1e-12)
            assert close(got, oracle.vertex_load(rp, col, incl), REL)
    with pytest.raises(ne.NecError):
        big = csr_from_adj([[] for _ in range(513)])
        ne.vertex_load_batch(ctx, [(big[0].astype(np.uint32), big[1])], True)


def test_markov_chain_measure_after_every_step(ctx, oracle):
    """BASELINE config 3 in miniature: chains of ErC graphs, one m_step, vertex_load(true) of all."""
    chains = [ne.ErEnsembleC(64, 4.0, 100 + g) for g in range(16)]
    for _ in range(3):
        for c in chains:
            c.m_step()
        csr = [c.export_csr() for c in chains]
        outs = ne.vertex_load_batch(ctx, [(rp.astype(np.uint32), col) for rp, col in csr], True)
        for (rp, col), got in zip(csr, outs):
            assert close(got, oracle.vertex_load(rp, col, True), REL)


def test_chains_step_and_load_in_cpp(ctx, oracle):
    chains = [ne.ErEnsembleC(48, 4.0, 500 + g) for g in range(12)] + [ne.SwEnsemble(40, 0.3, 9), ne.ErEnsembleM(30, 60, 4)]
    twins = [ne.ErEnsembleC(48, 4.0, 500 + g) for g in range(12)] + [ne.SwEnsemble(40, 0.3, 9), ne.ErEnsembleM(30, 60, 4)]
    for _ in range(2):
        outs = ne.chains_step_and_load(ctx, chains, 5, False)
        for t, got in zip(twins, outs):
            t.m_steps(5)
            rp, col = t.export_csr()
            assert close(got, oracle.vertex_load(rp, col, False), REL)


def test_invariants_at_scale(ctx):
    """Sizes the oracle cannot finish: integer invariants instead."""
    e = ne.SwEnsemble(1 << 14, 0.1, 123_239_010)
    dg = e.to_device(ctx)
    t, f = dg.vertex_load(True), dg.vertex_load(False)
    st = ctx.stats()
    n = 1 << 14
    assert st["reached_pairs"] == n * n                      # connected
    assert close(t - f, np.full(n, n - 1.0), REL)
    # sum_v load_true(v) = sum of all BFS distances; cross-check with a sampled set of sources
    src = np.arange(0, n, 257, dtype=np.uint32)
    part = dg.vertex_load_sources(src, True)
    dsum = sum(int(dg.bfs_debug(int(s))[0].astype(np.int64).sum()) for s in src[:8])
    part8 = dg.vertex_load_sources(src[:8], True)
    assert abs(part8.sum() - dsum) <= 1e-9 * dsum
    assert np.all(part <= t + 1e-6)
    dg.free()


def test_many_predecessors_and_hubs(engine_ctx, oracle):
    """A vertex with 300 predecessors (> 255: the mid engine hands the source back) and hubs of
    degree 300 (> 64: warp-cooperative expansion)."""
    ctx = engine_ctx
    edges = [(0, 1 + i) for i in range(300)] + [(1 + i, 301) for i in range(300)] + [(301, 302), (302, 303)]
    adj = adj_from_edges(304, edges)
    rp, col = csr_from_adj(adj)
    for incl in (True, False):
        assert close(_gpu_load(ctx, adj, incl), oracle.vertex_load(rp, col, incl), REL)
    b = ne.ba_scalable(5000, 4, 5, 11)
    rp, col = b.export_csr()
    assert max(np.diff(rp.astype(np.int64))) > 64
    dg = ctx.upload(rp, col)
    assert close(dg.vertex_load(True), oracle.vertex_load(rp, col, True, threads=0), REL)
    dg.free()


def test_batched_engine_reports_queue_overflow_loudly(ctx, oracle):
    """With a starved workspace the batched engine shrinks its level-queue bound from 32n to 8n entries per
    slot; a path graph puts every vertex on 32 distinct levels of a 32-source batch, which must be reported
    (NEC_EINTERNAL), never silently truncated — and the context must stay usable."""
    n = 6000
    rp, col = csr_from_adj(adj_from_edges(n, [(i, i + 1) for i in range(n - 1)]))
    dg = ctx.upload(rp, col)
    ctx.set_engine(1)
    ctx.set_workspace_limit(3 * 1000 * 1000)                    # one slot, and only with the shrunken queue
    with pytest.raises(ne.NecError) as ei:
        dg.vertex_load_sources(np.arange(64, dtype=np.uint32), True)
    assert ei.value.code == -5 and "queue overflow" in str(ei.value)
    ctx.set_workspace_limit(0)
    src = np.arange(0, n, 97, dtype=np.uint32)
    assert close(dg.vertex_load_sources(src, True), oracle.vertex_load(rp, col, True, sources=src), REL)
    ctx.set_engine(0)
    dg.free()


def test_mid_engine_queue_overflow_hands_sources_back(ctx, oracle, monkeypatch):
    """The mid engine's queue holds 2n entries (plain-store claims may append a vertex twice); a source
    that overflows it must be handed back to the batched engine, not expanded past the hole."""
    e = ne.SwEnsemble(3000, 0.2, 21)
    rp, col = e.export_csr()
    want = oracle.vertex_load(rp, col, True, threads=0)
    monkeypatch.setenv("NEC_MID_QUEUE_CAP", "1024")          # far below n: every source overflows
    dg = ctx.upload(rp, col)
    got = dg.vertex_load(True)
    st = ctx.stats()
    assert st["wide_batches"] == 3000 and (st["engine"] >> 4) == 2
    assert close(got, want, REL)
    monkeypatch.delenv("NEC_MID_QUEUE_CAP")
    assert close(dg.vertex_load(True), want, REL)
    assert ctx.stats()["engine"] == 2
    dg.free()


def test_mid_engine_is_the_default_for_mid_size_graphs(ctx, oracle):
    e = ne.SwEnsemble(4096, 0.1, 3)
    rp, col = e.export_csr()
    dg = ctx.upload(rp, col)
    got = dg.vertex_load(True)
    assert ctx.stats()["n_batches"] == 4096            # one source per CTA step, not 128 batches of 32
    assert close(got, oracle.vertex_load(rp, col, True, threads=0), REL)
    ctx.set_engine(1)
    got_b = dg.vertex_load(True)
    ctx.set_engine(0)
    assert ctx.stats()["n_batches"] == 128
    assert close(got, got_b, REL)
    dg.free()


def test_sampled_parity_on_larger_graph(ctx, oracle):
    g = ne.gnm_scalable(1 << 15, 1 << 17, 42)
    rp, col = g.export_csr()
    dg = ctx.upload(rp, col)
    src = np.arange(0, 1 << 15, 509, dtype=np.uint32)
    for incl in (True, False):
        assert close(dg.vertex_load_sources(src, incl), oracle.vertex_load(rp, col, incl, sources=src, threads=0), REL)
    b = ne.ba_scalable(1 << 14, 4, 5, 7)
    rp, col = b.export_csr()
    dg2 = ctx.upload(rp, col)
    src = np.arange(0, 1 << 14, 251, dtype=np.uint32)
    assert close(dg2.vertex_load_sources(src, True), oracle.vertex_load(rp, col, True, sources=src, threads=0), REL)
    dg.free(), dg2.free()


def _random_graph(rng, n, m):
    """Simple undirected graph with ~m distinct edges, adjacency in insertion order."""
    adj = [[] for _ in range(n)]
    seen = set()
    if n < 2:
        return adj
    for _ in range(m):
        a, b = int(rng.integers(n)), int(rng.integers(n))
        if a == b or (min(a, b), max(a, b)) in seen:
            continue
        seen.add((min(a, b), max(a, b)))
        adj[a].append(b)
        adj[b].append(a)
    return adj


def test_differential_fuzz_all_engines(ctx, oracle):
    """Random small graphs of every density (isolated vertices, many components, near-complete), both
    traversal engines, both endpoint conventions, the small-graph batch kernel and the distance measures."""
    rng = np.random.default_rng(20261019)
    graphs = []
    for trial in range(120):
        n = int(rng.integers(1, 220))
        dens = float(rng.choice([0.0, 0.3, 0.6, 1.0, 2.0, 4.0, 12.0]))
        graphs.append(_random_graph(rng, n, int(dens * n)))
    csr = [csr_from_adj(a) for a in graphs]
    want_t = [oracle.vertex_load(rp, col, True) for rp, col in csr]
    want_f = [oracle.vertex_load(rp, col, False) for rp, col in csr]
    for engine, width in ((1, "32"), (1, "64"), (2, "32")):
        ctx.set_engine(engine)
        os.environ["NEC_BATCH_WIDTH"] = width
        for (rp, col), wt, wf in zip(csr, want_t, want_f):
            dg = ctx.upload(rp, col)
            assert close(dg.vertex_load(True), wt, REL) and close(dg.vertex_load(False), wf, REL)
            ecc, sumd, reached = dg.distance_measures()
            recc, rsum, rreach = oracle.distance_measures(rp, col)
            assert ecc.tolist() == recc.tolist() and sumd.tolist() == rsum.tolist() and reached.tolist() == rreach.tolist()
            dg.free()
    ctx.set_engine(0)
    del os.environ["NEC_BATCH_WIDTH"]
    outs = ne.vertex_load_batch(ctx, [(rp.astype(np.uint32), col) for rp, col in csr], False)
    for got, wf in zip(outs, want_f):
        assert close(got, wf, REL)


def test_mid_engine_hub_list_overflow_and_hub_heavy_graphs(ctx, oracle):
    """600 vertices of degree 71 on one BFS level: more hubs than the shared hub list holds (512), so some
    are expanded inline; plus a root of degree 600 and a Barabasi-Albert graph with hubs of degree ~1000."""
    edges = [(0, 1 + h) for h in range(600)]
    nxt = 601
    for h in range(600):
        for _ in range(70):
            edges.append((1 + h, nxt))
            nxt += 1
    adj = adj_from_edges(nxt, edges)
    rp, col = csr_from_adj(adj)
    dg = ctx.upload(rp, col)
    src = np.concatenate([np.arange(0, 8), np.arange(601, nxt, 997)]).astype(np.uint32)
    ctx.set_engine(2)
    for incl in (True, False):
        assert close(dg.vertex_load_sources(src, incl), oracle.vertex_load(rp, col, incl, sources=src), REL)
    assert ctx.stats()["engine"] == 2
    ctx.set_engine(0)
    dg.free()
    b = ne.ba_scalable(60000, 4, 5, 3)
    rp, col = b.export_csr()
    assert int(np.diff(rp.astype(np.int64)).max()) > 500
    dg = ctx.upload(rp, col)
    src = np.arange(0, 60000, 1201, dtype=np.uint32)
    assert close(dg.vertex_load_sources(src, True), oracle.vertex_load(rp, col, True, sources=src, threads=0), REL)
    assert ctx.stats()["engine"] & 15 in (1, 2)
    dg.free()


def test_wide_batches_are_the_default_for_many_sources(ctx, oracle):
    """>= 64 sources per SM's worth of batches: the batched engine runs 64 sources per batch (two per
    lane); a ragged tail batch and sources that repeat are covered."""
    g = ne.ErEnsembleC(3000, 3.0, 11)                        # many components, isolated vertices
    rp, col = g.export_csr()
    dg = ctx.upload(rp, col)
    ctx.set_engine(1)
    rng = np.random.default_rng(5)
    src = rng.integers(0, 3000, 64 * 148 + 37).astype(np.uint32)   # with repeats; 37-source tail
    got = dg.vertex_load_sources(src, True)
    st = ctx.stats()
    assert st["batch_width"] == 64 and st["n_batches"] == 149
    assert close(got, oracle.vertex_load(rp, col, True, sources=src, threads=0), REL)
    ecc, sumd, reached = dg.distance_measures(src)
    assert ctx.stats()["batch_width"] == 64
    recc, rsum, rreach = oracle.distance_measures(rp, col, src)
    assert ecc.tolist() == recc.tolist() and sumd.tolist() == rsum.tolist() and reached.tolist() == rreach.tolist()
    got = dg.vertex_load_sources(src[:2000], True)
    assert ctx.stats()["batch_width"] == 32
    assert close(got, oracle.vertex_load(rp, col, True, sources=src[:2000], threads=0), REL)
    ctx.set_engine(0)
    dg.free()


def test_chain_batch_pipeline_many_chunks(ctx, oracle):
    """> 256 graphs per call: the batch call runs as a pipeline of chunks (pack -> copy in -> kernel -> copy
    out); mixed sizes, uniform sizes (2-D result), and a malformed graph in a late chunk."""
    rng = np.random.default_rng(3)
    sizes = rng.integers(20, 70, 1100)
    chains = [ne.ErEnsembleC(int(n), 3.0, 9000 + g) for g, n in enumerate(sizes)]
    twins = [ne.ErEnsembleC(int(n), 3.0, 9000 + g) for g, n in enumerate(sizes)]
    batch = ne.ChainBatch(chains)
    for step in range(2):
        outs = batch.step_and_load(ctx, 3, True)
        assert batch.stats["kernel_launches"] == 4 + 1 and batch.stats["engine"] == 3    # 1100 graphs -> 4 chunks, + the replay of the step's list operations
        assert batch.stats["n_sources"] == int(sizes.sum())
        for t in twins:
            t.m_steps(3)
        for g in list(range(0, 1100, 37)) + [1099]:
            rp, col = twins[g].export_csr()
            assert close(outs[g], oracle.vertex_load(rp, col, True), REL), g
    uni = ne.ChainBatch([ne.ErEnsembleC(40, 4.0, 70 + g) for g in range(600)])
    res = uni.step_and_load(ctx, 0, False)
    assert res.shape == (600, 40)
    rp, col = uni.chains[599].export_csr()
    assert close(res[599], oracle.vertex_load(rp, col, False), REL)
    # malformed input in the last chunk: loud error, nothing half-written, context still usable
    graphs = [c.export_csr() for c in chains[:700]]
    graphs = [(rp.astype(np.uint32), col.copy()) for rp, col in graphs]
    bad = next(g for g in range(650, 700) if len(graphs[g][1]))
    graphs[bad][1][0] = 10_000
    with pytest.raises(ne.NecError, match="out of range"):
        ne.vertex_load_batch(ctx, graphs, True)
    graphs[bad][1][0] = 0
    outs = ne.vertex_load_batch(ctx, graphs[:600], True)
    assert ctx.stats()["kernel_launches"] == 2
    for g in (0, 299, 300, 599):
        rp, col = chains[g].export_csr()
        assert close(outs[g], oracle.vertex_load(rp, col, True), REL)


def _full_size_properties(ctx, oracle, rp, col, src, want_engine, want_width):
    """Size-independent properties at BASELINE's full sizes (the oracle only sees a handful of sources there):
    with S the source set, R_s the vertices s reaches and d_s the BFS distances,
      sum_v (load_true - load_false)(v) = sum_{s in S} (|R_s| - 1)       (an endpoint adds exactly 1),
      sum_v load_true(v)               = sum_{s in S} sum_v d_s(v)       (a unit from v crosses d_s(v) vertices),
    the right-hand sides being the exact integers of nec_distance_measures (forward pass alone)."""
    dg = ctx.upload(rp, col)
    if callable(src):
        src = src(dg)
    lt = dg.vertex_load_sources(src, True)
    st = ctx.stats()
    assert st["engine"] & 15 == want_engine and (want_width is None or st["batch_width"] == want_width)
    lf = dg.vertex_load_sources(src, False)
    ecc, sumd, reached = dg.distance_measures(src)
    assert st["reached_pairs"] == int(reached.astype(np.int64).sum())
    ends = int((reached.astype(np.int64) - 1).sum())
    assert abs(float((lt - lf).sum()) - ends) <= 1e-9 * ends
    dtot = int(sumd.astype(object).sum())
    assert abs(float(lt.sum()) - dtot) <= 1e-9 * dtot
    diff = lt - lf                                          # per vertex: how many sources other than v reach it
    assert np.all(np.abs(diff - np.rint(diff)) <= 1e-9 * np.maximum(1.0, diff)) and diff.min() >= -1e-6 and diff.max() <= len(src) * (1 + 1e-9)
    sub = src[:: max(1, len(src) // 6)][:6]
    assert close(dg.vertex_load_sources(sub, True), oracle.vertex_load(rp, col, True, sources=sub, threads=0), REL)
    dg.free()
    return st


def test_full_size_properties_config2_small_world(ctx, oracle):
    e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)          # BASELINE configs[1], every source
    rp, col = e.export_csr()
    st = _full_size_properties(ctx, oracle, rp, col, np.arange(1 << 16, dtype=np.uint32), 2, None)
    assert st["reached_pairs"] == (1 << 32)                  # connected


def test_full_size_properties_config4_gnm(ctx, oracle):
    """BASELINE configs[3] (the north-star target) on the path the bench takes: one resident wave through the DEFAULT
    plan = the compact engine at its widest batch (256 sources, entries with more than 128 live sources consumed in
    two passes) on clusters of 1024-thread CTAs that fill the GPU; then the 128-wide shape on the same sources."""
    g = ne.gnm_scalable(1 << 20, 1 << 22, 20261020)          # (scalable generator), sampled sources
    rp, col = g.export_csr()

    def one_wave(dg):
        wave = dg.wave_size()
        assert wave % 128 == 0 and wave >= 9000
        return (np.arange(wave, dtype=np.uint64) * 110_503 % (1 << 20)).astype(np.uint32)

    st = _full_size_properties(ctx, oracle, rp, col, one_wave, 4, None)
    assert st["batch_width"] in (128, 256) and st["cta_threads"] == 1024 and st["pull_levels"] > 0
    assert st["n_slots"] * st["batch_width"] == st["n_sources"] and 130 <= st["n_slots"] * st["team_size"] <= 148
    assert st["requeued_batches"] == 0 and st["n_sources"] * ((1 << 20) - 2000) < st["reached_pairs"] <= st["n_sources"] * (1 << 20)   # a few hundred isolated vertices
    # the plan must not depend on how many waves a call carries: a two-wave call gets the same shape
    dg = ctx.upload(rp, col)
    wave = dg.wave_size()
    two = (np.arange(2 * wave, dtype=np.uint64) * 110_503 % (1 << 20)).astype(np.uint32)
    both = dg.vertex_load_sources(two, True)
    st2 = ctx.stats()
    assert (st2["batch_width"], st2["team_size"], st2["n_slots"], st2["n_batches"]) == (st["batch_width"], st["team_size"], st["n_slots"], 2 * st["n_slots"])
    first = dg.vertex_load_sources(two[:wave], True)
    second = dg.vertex_load_sources(two[wave:], True)
    assert close(both, first + second, 1e-12)
    dg.free()


@pytest.mark.parametrize("width,team", [("128", 2), ("256", 3)])
def test_config4_batch_widths_agree(ctx, oracle, monkeypatch, width, team):
    """Both production shapes of the compact engine on config 4 against the oracle on the same sampled sources."""
    g = ne.gnm_scalable(1 << 20, 1 << 22, 20261020)
    rp, col = g.export_csr()
    monkeypatch.setenv("NEC_WIDE_WIDTH", width)
    dg = ctx.upload(rp, col)
    src = (np.arange(1536, dtype=np.uint64) * 110_503 % (1 << 20)).astype(np.uint32)
    got = dg.vertex_load_sources(src, True)
    st = ctx.stats()
    assert st["engine"] == 4 and st["batch_width"] == int(width)
    sub = src[::192]
    assert close(dg.vertex_load_sources(sub, True), oracle.vertex_load(rp, col, True, sources=sub, threads=0), REL)
    ecc, sumd, reached = dg.distance_measures(src)
    dtot = int(sumd.astype(object).sum())
    assert abs(float(got.sum()) - dtot) <= 1e-9 * dtot and st["reached_pairs"] == int(reached.astype(np.int64).sum())
    dg.free()


def test_full_size_properties_config5_barabasi_albert(ctx, oracle):
    """BASELINE configs[4] on the path the bench takes: one resident wave of sources through the DEFAULT plan, which
    on this graph is the compact engine on few (memory-heavy) slots, each run by a team of many CTAs."""
    g = ne.ba_scalable(1 << 22, 4, 5, 20261020)              # (scalable generator), sampled sources
    rp, col = g.export_csr()

    def one_wave(dg):
        wave = dg.wave_size()
        assert wave % 64 == 0 and wave >= 2000                 # (large teams on few slots: 9 teams of 16 CTAs x 256 sources)
        return (np.arange(wave, dtype=np.uint64) * 2_654_435_761 % (1 << 22)).astype(np.uint32)

    st = _full_size_properties(ctx, oracle, rp, col, one_wave, 4, None)
    assert st["batch_width"] in (64, 128, 256) and st["team_size"] >= 3 and st["n_slots"] * st["batch_width"] == st["n_sources"]
    assert st["n_slots"] * st["team_size"] <= 148 * (2 if st["cta_threads"] == 512 else 1)
    assert st["pull_levels"] > 0 and st["push_levels"] > 0    # direction chosen per level
    assert st["requeued_batches"] == 0 and st["reached_pairs"] == st["n_sources"] * (1 << 22)  # a BA graph is connected


def test_config5_small_calls(ctx, oracle):
    g = ne.ba_scalable(1 << 22, 4, 5, 20261020)
    rp, col = g.export_csr()
    src = (np.arange(1536, dtype=np.uint64) * 2_654_435_761 % (1 << 22)).astype(np.uint32)
    st = _full_size_properties(ctx, oracle, rp, col, src, 4, None)   # a dozen batches: large teams on few slots
    assert st["reached_pairs"] == 1536 * (1 << 22) and st["n_slots"] == st["n_batches"] and st["team_size"] >= 4
    ctx.set_engine(1)                                               # the source-minor engine at this size: narrow batches
    st = _full_size_properties(ctx, oracle, rp, col, src, 1, 32)
    ctx.set_engine(0)
    assert st["reached_pairs"] == 1536 * (1 << 22)


def test_config3_verbatim_4096_chains_of_256(ctx, oracle):
    """BASELINE configs[2] at its real size: 4096 Markov chains of ErC N=256 c=4 in one ChainBatch, three
    measurement steps; 96 sampled chains are compared with the oracle on independently stepped twins."""
    seeds = [123_239_010 + g for g in range(4096)]
    batch = ne.ChainBatch([ne.ErEnsembleC(256, 4.0, s) for s in seeds])
    pick = list(range(0, 4096, 43))[:96] + [4095]
    twins = {g: ne.ErEnsembleC(256, 4.0, seeds[g]) for g in pick}
    for step in range(3):
        outs = batch.step_and_load(ctx, 1, True)
        assert outs.shape == (4096, 256) and batch.stats["n_sources"] == 4096 * 256 and batch.stats["engine"] == 3
        for g in pick:
            twins[g].m_steps(1)
            rp, col = twins[g].export_csr()
            assert close(outs[g], oracle.vertex_load(rp, col, True), REL), (step, g)


@pytest.mark.parametrize("width", ["32", "64"])
def test_cluster_teams_are_bitwise_identical_to_single_ctas(ctx, oracle, monkeypatch, width):
    """Memory-limited slot counts run each slot on a thread-block cluster of 512-thread CTAs (shared queue tail
    in distributed shared memory, cluster barriers).  Same writers, same summation order: bit-identical."""
    g = ne.ErEnsembleC(1800, 3.0, 77)                        # several components, isolated vertices
    rp, col = g.export_csr()
    ba = ne.ba_scalable(5000, 3, 4, 9)                       # hubs
    rp2, col2 = ba.export_csr()
    ctx.set_engine(1)
    monkeypatch.setenv("NEC_BATCH_WIDTH", width)
    monkeypatch.setenv("NEC_BATCHED_WIDE", "1")
    for rp_, col_, src in ((rp, col, None), (rp2, col2, np.arange(0, 5000, 3, dtype=np.uint32))):
        dg = ctx.upload(rp_, col_)
        monkeypatch.setenv("NEC_TEAM_SIZE", "1")
        one = dg.vertex_load(False) if src is None else dg.vertex_load_sources(src, False)
        monkeypatch.setenv("NEC_TEAM_SIZE", "3")
        team = dg.vertex_load(False) if src is None else dg.vertex_load_sources(src, False)
        st = ctx.stats()
        monkeypatch.setenv("NEC_TEAM_SIZE", "4")
        team4 = dg.vertex_load(False) if src is None else dg.vertex_load_sources(src, False)
        assert one.tobytes() == team.tobytes() == team4.tobytes()
        assert close(team, oracle.vertex_load(rp_, col_, False, sources=src, threads=0), REL)
        assert st["batch_width"] == int(width)
        # few slots, many batches per slot: the team re-uses queues, masks and rows batch after batch
        # (partial loads are summed per slot, so the comparison is at the same slot count)
        # (only on the shallow BA graph: a tight budget also shrinks the queues to 8 entries per vertex)
        if rp_ is rp:
            dg.free()
            continue
        ctx.set_workspace_limit(16_000_000 * int(width) // 32)
        monkeypatch.setenv("NEC_TEAM_SIZE", "1")
        few = dg.vertex_load(False) if src is None else dg.vertex_load_sources(src, False)
        monkeypatch.setenv("NEC_TEAM_SIZE", "3")
        again = dg.vertex_load(False) if src is None else dg.vertex_load_sources(src, False)
        st = ctx.stats()
        assert st["n_batches"] >= 3 * st["n_slots"] and again.tobytes() == few.tobytes() and close(again, one, 1e-12)
        ctx.set_workspace_limit(0)
        dg.free()
    ctx.set_engine(0)


def test_wave_size_matches_what_a_call_uses(ctx):
    sw = ne.SwEnsemble(4096, 0.1, 3)
    dg = sw.to_device(ctx)
    wave = dg.wave_size()
    assert wave in (296, 1184) and dg.wave_size(10) == 10             # mid engine: one source per CTA; 2 x 512 or (small graphs) 8 x 128 threads per SM
    dg.vertex_load_sources(np.arange(wave, dtype=np.uint32), True)
    assert ctx.stats()["n_slots"] == wave and ctx.stats()["cta_threads"] == (128 if wave == 1184 else 512)
    dg.free()
    g = ne.gnm_scalable(200_000, 800_000, 5)                           # beyond the mid engine: compact batched engine
    rp, col = g.export_csr()
    dg = ctx.upload(rp, col)
    wave = dg.wave_size()
    src = np.arange(wave, dtype=np.uint32)
    dg.vertex_load_sources(src, True)
    st = ctx.stats()
    assert st["engine"] == 4 and st["n_slots"] * st["batch_width"] == wave and st["n_batches"] == st["n_slots"]
    ctx.set_engine(1)                                                  # and the source-minor one, when asked for
    wave = dg.wave_size()
    dg.vertex_load_sources(np.arange(wave, dtype=np.uint32), True)
    st = ctx.stats()
    assert st["engine"] == 1 and st["n_slots"] * st["batch_width"] == wave and st["n_batches"] == st["n_slots"]
    ctx.set_engine(0)
    dg.free()


@pytest.mark.parametrize("width", ["32", "64"])
def test_forward_directions_are_bitwise_identical(ctx, oracle, monkeypatch, width):
    """The batched forward pass expands a level by push (frontier -> neighbours, atomics) or by pull (every
    incomplete vertex gathers its neighbours' frontier masks), chosen per level by a cost model.  Level sets and
    masks are the same either way and no sum depends on the order of a level's queue: bit-identical loads."""
    ctx.set_engine(1)
    monkeypatch.setenv("NEC_BATCH_WIDTH", width)
    graphs = [ne.gnm_scalable(30000, 120000, 3).export_csr(), ne.ba_scalable(20000, 4, 5, 5).export_csr(),
              ne.ErEnsembleC(2500, 1.2, 8).export_csr()]                     # expander, hubs, many small components
    for rp, col in graphs:
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        src = np.arange(0, n, max(1, n // 200), dtype=np.uint32)
        outs, stats = [], []
        for mode in ("0", "1", "2"):
            monkeypatch.setenv("NEC_PULL", mode)
            outs.append(dg.vertex_load_sources(src, False))
            stats.append(ctx.stats())
        assert outs[0].tobytes() == outs[1].tobytes() == outs[2].tobytes()
        assert stats[0]["pull_levels"] == 0 and stats[0]["push_levels"] > 0
        assert stats[2]["push_levels"] == 0 and stats[2]["pull_levels"] == stats[0]["push_levels"]
        assert stats[1]["pull_levels"] + stats[1]["push_levels"] == stats[0]["push_levels"]
        assert [s["reached_pairs"] for s in stats] == [stats[0]["reached_pairs"]] * 3
        assert [s["vertex_visits"] for s in stats] == [stats[0]["vertex_visits"]] * 3
        assert close(outs[1], oracle.vertex_load(rp, col, False, sources=src, threads=0), REL)
        ecc, sumd, reached = dg.distance_measures(src)                      # the forward pass alone, pulled
        recc, rsum, rreach = oracle.distance_measures(rp, col, src)
        assert ecc.tolist() == recc.tolist() and sumd.tolist() == rsum.tolist() and reached.tolist() == rreach.tolist()
        dg.free()
    ctx.set_engine(0)


def test_device_resident_chains_replay_list_operations_exactly(ctx, oracle):
    """the chains' graphs stay on the device; a step ships the list operations the host chains
    performed (push / swap_remove / retarget) and the device replays them.  After 1000 steps the device adjacency
    ORDER equals the host twins' exactly, for every ensemble kind with a Markov chain, and the loads match the
    oracle; one upload in total."""
    def make():
        return ([ne.ErEnsembleC(256, 4.0, 123_239_010 + g) for g in range(40)] + [ne.ErEnsembleC(50, 3.0, 9 + g) for g in range(10)] +
                [ne.SwEnsemble(100, 0.15, 31 + g) for g in range(10)] + [ne.ErEnsembleM(40, 90, 5 + g) for g in range(6)])
    chains, twins = make(), make()
    batch = ne.ChainBatch(chains)
    assert batch.device_resident
    for rounds, steps in ((3, 1), (1, 997)):
        for _ in range(rounds):
            outs = batch.step_and_load(ctx, steps, True)
            for t in twins:
                t.m_steps(steps)
    assert batch.uploads == 1 and batch.h2d_bytes_last_step < 64 * 1024 * 8
    dev_adj = batch.device_adjacency()
    for g, t in enumerate(twins):
        assert dev_adj[g] == t.adjacency(), g                       # order included
        assert chains[g].adjacency() == t.adjacency()
        rp, col = t.export_csr()
        assert close(outs[g], oracle.vertex_load(rp, col, True), REL), g
    # a step outside the batch (or a randomize) is noticed: the device copy is rebuilt, nothing is silently stale
    chains[0].m_steps(5); twins[0].m_steps(5)
    chains[1].randomize(); twins[1].randomize()
    outs = batch.step_and_load(ctx, 0, False)
    assert batch.uploads == 2
    for g in (0, 1, 2):
        rp, col = twins[g].export_csr()
        assert close(outs[g], oracle.vertex_load(rp, col, False), REL)
    # same numbers as the re-export path, bit for bit (same kernel arithmetic on the same lists)
    plain = ne.ChainBatch(chains, device_resident=False).step_and_load(ctx, 0, False)
    for g in range(len(chains)):
        assert np.asarray(outs[g]).tobytes() == np.asarray(plain[g]).tobytes()
    batch.close()


def test_device_resident_chains_grow_their_rows(ctx, oracle):
    """Lists that outgrow the padded rows: the device copy is rebuilt with more slack, results stay right."""
    chains = [ne.ErEnsembleC(64, 1.0, 3 + g) for g in range(8)]
    batch = ne.ChainBatch(chains)
    batch.step_and_load(ctx, 0, True)
    for c in chains:                       # densify far beyond the initial degrees, behind the op log via the batch
        pass
    g0 = chains[0]
    for a in range(1, 40):                 # vertex 0 gets degree 39: logged as wholesale-free list operations
        g0.add_edge(0, a)
    outs = batch.step_and_load(ctx, 0, True)
    rp, col = g0.export_csr()
    assert close(outs[0], oracle.vertex_load(rp, col, True), REL)
    assert batch.device_adjacency()[0] == g0.adjacency()
    batch.close()


def test_erc_randomize_on_device_matches_the_fixtures_bit_for_bit(ctx, oracle, golden):
    """ErEnsembleC::randomize drawn on the device by Pcg64 jump-ahead.  From the fixture's
    seed the reference's constructor is new() = one randomize; the device sample must reproduce the fixture's
    adjacency lists (order included) and final generator state, and keep doing so sample after sample."""
    for name in ("erc_1", "erc_2"):
        fx = golden["graphs"][name]
        n, c, seed = fx["n"], fx["c"], fx["seed"]
        host = ne.ErEnsembleC(n, c, seed)                      # host loop: equals the fixture (tests/test_generators.py)
        assert host.adjacency() == fx["adj"]
        # a second ensemble whose FIRST randomize we redo on the device from the seed state
        lib = ne._lib.load_host() if hasattr(ne, "_lib") else None
        dev = ne.ErEnsembleC(n, c, seed)
        # rewind: construct the seed state by hand is not exposed, so compare sample k+1 instead: both ensembles
        # sit after the constructor's randomize; the next sample comes from the host loop on one, the device on the other
        for _ in range(3):
            host.randomize()
            dg = dev.randomize_cuda(ctx)
            assert dev.adjacency() == host.adjacency()
            assert dev.rng_state() == host.rng_state()
            rp, col = dg.download()
            hrp, hcol = host.export_csr()
            assert rp.tolist() == hrp.tolist() and col.tolist() == hcol.tolist()
            assert close(dg.vertex_load(True), oracle.vertex_load(hrp, hcol, True), REL)
            dg.free()
    # the fixture itself: state after the constructor == fixture's recorded generator, so the chain of equalities
    # above starts from a pinned point
    fx = golden["graphs"]["erc_1"]
    assert ne.ErEnsembleC(fx["n"], fx["c"], fx["seed"]).rng_state() == (int(fx["rng_state"]), int(fx["rng_increment"]))
    # larger and denser samples (rows longer than a warp's 32 pairs; lower parts that need sorting)
    for n, c, seed in ((1000, 4.0, 123_239_010), (700, 40.0, 5), (64, 63.0, 2), (2, 1.0, 1)):
        a, b = ne.ErEnsembleC(n, c, seed), ne.ErEnsembleC(n, c, seed)
        a.randomize()
        dg = b.randomize_cuda(ctx)
        assert a.adjacency() == b.adjacency() and a.rng_state() == b.rng_state()
        assert dg.n == n and dg.nnz == 2 * a.edge_count()
        dg.free()


def test_path_counts_exact_on_larger_graphs(ctx, oracle):
    """f64 shortest-path counts (sigma) from the engine's own level queues (pull, lane per source), exact while
    they fit in f64 - on graphs far beyond the fixtures: an expander, a hub-heavy BA graph, a grid whose counts are
    binomials (large but < 2^53), and a deep path (32-bit distance rows).  sigma never enters the load."""
    side = 26                                                      # C(50, 25) ~ 1.26e14 < 2^53
    grid = adj_from_edges(side * side, [(r * side + c, r * side + c + 1) for r in range(side) for c in range(side - 1)] +
                          [(r * side + c, (r + 1) * side + c) for r in range(side - 1) for c in range(side)])
    cases = [ne.gnm_scalable(1 << 15, 1 << 17, 9).export_csr(), ne.ba_scalable(20000, 4, 5, 2).export_csr(), csr_from_adj(grid),
             csr_from_adj(adj_from_edges(400, [(i, i + 1) for i in range(399)]))]
    for rp, col in cases:
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        for s in (0, n // 2, n - 1):
            dist, npred, sigma, bk = dg.bfs_debug(s)
            rd, rn, rs, rb = oracle.bfs_debug(rp, col, s)
            assert dist.tolist() == rd.tolist() and npred.tolist() == rn.tolist()
            assert sigma.tolist() == rs.tolist()                  # bit-exact
            assert close(bk, rb, REL)
        dg.free()
    import math
    dg = ctx.upload(*csr_from_adj(grid))
    assert dg.bfs_debug(0)[2][side * side - 1] == float(math.comb(2 * (side - 1), side - 1))
    dg.free()


def test_mid_engine_cta_shapes_agree(ctx, oracle, monkeypatch):
    """Small graphs run the one-source-per-CTA engine with eight 128-thread CTAs per SM, larger ones with two
    512-thread CTAs; same arithmetic, same per-source order: identical per-source values, loads equal to rounding
    of the slot sums."""
    for e in (ne.ErEnsembleC(1000, 4.0, 123_239_010), ne.SwEnsemble(3000, 0.1, 4), ne.ba_scalable(4000, 3, 4, 6)):
        rp, col = e.export_csr()
        dg = ctx.upload(rp, col)
        monkeypatch.setenv("NEC_MID_SMALL_N", "0")
        big = dg.vertex_load(False)
        assert ctx.stats()["cta_threads"] == 512 and ctx.stats()["engine"] & 15 == 2
        monkeypatch.setenv("NEC_MID_SMALL_N", "100000")
        small = dg.vertex_load(False)
        assert ctx.stats()["cta_threads"] == 128
        monkeypatch.delenv("NEC_MID_SMALL_N")
        assert close(small, big, 1e-12) and close(small, oracle.vertex_load(rp, col, False, threads=0), REL)
        dg.free()


def test_mid_engine_consecutive_sources_need_no_list_and_small_graphs_reduce_in_groups(ctx, oracle):
    """A run of consecutive source ids is passed to the kernel as (first id, count) instead of a device list, and on small
    graphs the slots' partial loads are summed by slot groups (many slots, few vertices): same values as the list path
    and the oracle, whatever the first id, and bitwise reproducible."""
    for e in (ne.ErEnsembleC(1000, 4.0, 5), ne.SwEnsemble(3000, 0.1, 9)):
        rp, col = e.export_csr()
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        ctx.set_engine(2)
        for lo, hi in ((0, n), (n // 3, n // 3 + 257), (n - 40, n), (17, 18)):
            run = np.arange(lo, hi, dtype=np.uint32)
            got = dg.vertex_load_sources(run, True)
            assert ctx.stats()["engine"] == 2 and ctx.stats()["n_sources"] == hi - lo
            assert dg.vertex_load_sources(run, True).tobytes() == got.tobytes()
            rev = dg.vertex_load_sources(run[::-1].copy(), True)             # same sources, not a run: the list path
            assert close(got, rev, 1e-12)
            assert close(got, oracle.vertex_load(rp, col, True, sources=run, threads=0), REL), (lo, hi)
        ctx.set_engine(0)
        dg.free()


def test_engine_is_chosen_by_structure_for_graphs_that_fit_the_mid_engine(ctx, oracle, monkeypatch):
    """From n = 8192 up a large call on a graph that fits the mid engine first probes one batch with the compact batched engine
    (queue entries per vertex: how many distinct levels a vertex sits on within a batch) and keeps the finding with the
    graph: random / scale-free graphs go to the compact engine, small-world rings stay with one source per CTA.  The
    choice is a function of the graph alone: repeated calls are bitwise equal, small calls never probe, a forced
    engine is obeyed, and all of them agree with the oracle."""
    cases = [(ne.gnm_scalable(20000, 80000, 3), 4), (ne.ba_scalable(12000, 3, 4, 7), 4), (ne.SwEnsemble(12000, 0.05, 9), 2)]
    for e, want in cases:
        rp, col = e.export_csr()
        n = len(rp) - 1
        dg = ctx.upload(rp, col)
        few = np.arange(0, n, n // 100, dtype=np.uint32)                  # a small call: no probe, mid engine
        got_few = dg.vertex_load_sources(few, True)
        assert ctx.stats()["engine"] == 2
        assert close(got_few, oracle.vertex_load(rp, col, True, sources=few, threads=0), REL)
        src = np.arange(0, n, 4, dtype=np.uint32)                         # a large call
        got = dg.vertex_load_sources(src, True)
        assert ctx.stats()["engine"] & 15 == want, (n, want, ctx.stats())
        assert dg.vertex_load_sources(src, True).tobytes() == got.tobytes()
        wave = dg.wave_size(len(src))
        if want == 2:
            assert wave in (min(len(src), 296), min(len(src), 1184))
        else:
            assert wave == min(len(src), ctx.stats()["n_slots"] * ctx.stats()["batch_width"])
        want_load = oracle.vertex_load(rp, col, True, sources=src, threads=0)
        assert close(got, want_load, REL)
        ctx.set_engine(2)
        forced = dg.vertex_load_sources(src, True)
        assert ctx.stats()["engine"] == 2 and close(forced, want_load, REL) and close(forced, got, 1e-12)
        ctx.set_engine(0)
        monkeypatch.setenv("NEC_PROBE", "0")                                # size threshold only
        dg.vertex_load_sources(src, True)
        assert ctx.stats()["engine"] == 2
        monkeypatch.delenv("NEC_PROBE")
        dg.free()


def test_smoke_entry_point_runs():
    """The driver's smoke(): kept under test so that a change of defaults cannot break it unnoticed."""
    import __graft_entry__
    __graft_entry__.smoke()
