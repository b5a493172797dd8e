"""Multi-GPU behind the C ABI: one context over several devices (nec_init with n_devices > 1) — worker thread,
stream and arena per device, replicated graph, sources dealt in 64-blocks, one ncclAllReduce.  The reference's
single call (src/traits/graph_traits.rs:328-330) then runs on every GPU of the box.  Needs >= 2 GPUs: skipped on
a one-GPU box (the sharding arithmetic itself is covered on CPU by tests/test_distributed.py)."""
import numpy as np
import pytest

import net_ensembles_b200 as ne
from tests.helpers import close

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def ctx2(built):
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    c = ne.Context(devices=[0, 1])
    yield c
    c.close()


def test_vertex_load_on_two_devices_matches_one_device_and_oracle(ctx2, ctx, oracle):
    for e in (ne.SwEnsemble(6000, 0.1, 5), ne.ErEnsembleC(1000, 4.0, 123_239_010), ne.gnm_scalable(200_000, 800_000, 3)):
        rp, col = e.export_csr()
        n = len(rp) - 1
        dg2, dg1 = e.to_device(ctx2), ctx.upload(rp, col)
        src = None if n <= 6000 else np.arange(0, n, n // 640, dtype=np.uint32)[:640]
        for incl in (True, False):
            two = dg2.vertex_load(incl) if src is None else dg2.vertex_load_sources(src, incl)
            st2 = ctx2.stats()
            one = dg1.vertex_load(incl) if src is None else dg1.vertex_load_sources(src, incl)
            st1 = ctx.stats()
            assert close(two, one, 1e-12)
            assert st2["reached_pairs"] == st1["reached_pairs"] and st2["edge_pairs"] == st1["edge_pairs"]
            assert st2["n_sources"] == st1["n_sources"]
        want = oracle.vertex_load(rp, col, True, sources=src, threads=0)
        got = dg2.vertex_load(True) if src is None else dg2.vertex_load_sources(src, True)
        assert close(got, want, 1e-9)
        again = dg2.vertex_load(True) if src is None else dg2.vertex_load_sources(src, True)
        assert got.tobytes() == again.tobytes()                    # reproducible run to run
        dg2.free(), dg1.free()


def test_drop_in_call_batch_and_measures_on_two_devices(ctx2, oracle):
    e = ne.ErEnsembleC(700, 3.0, 17)
    rp, col = e.export_csr()
    assert close(e.vertex_load(False, ctx2), oracle.vertex_load(rp, col, False), 1e-9)      # the ensemble's own call
    dg = e.to_device(ctx2)
    ecc, sumd, reached = dg.distance_measures()
    recc, rsum, rreach = oracle.distance_measures(rp, col)
    assert ecc.tolist() == recc.tolist() and sumd.tolist() == rsum.tolist() and reached.tolist() == rreach.tolist()
    assert dg.wave_size() in (2 * 296, 700)
    dg.free()
    chains = [ne.ErEnsembleC(64, 4.0, 7 + g).export_csr() for g in range(37)]
    outs = ne.vertex_load_batch(ctx2, [(r.astype(np.uint32), c) for r, c in chains], True)
    for (r, c), o in zip(chains, outs):
        assert close(o, oracle.vertex_load(r, c, True), 1e-9)
    with pytest.raises(ne.NecError):
        ctx2.set_stream(12345)


def test_more_devices_than_the_box_has_is_refused(built):
    n = _device_count()
    with pytest.raises(ne.NecError):
        ne.Context(devices=list(range(n + 1)))
    with pytest.raises(ne.NecError):
        ne.Context(devices=[0, 0])
