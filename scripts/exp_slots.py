"""Experiment: per-batch time as a function of resident slots (footprint / TLB reach)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import net_ensembles_b200 as ne
e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
rp, col = e.export_csr()
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
for gb in (0.3, 0.6, 1.2, 2.4, 4.8, 9.6, 20):
    ctx.set_workspace_limit(int(gb * 1e9))
    # probe slots
    dg.vertex_load_sources(np.arange(32 * 400, dtype=np.uint32), True)
    slots = ctx.stats()["n_slots"]
    nb = 2 * slots
    src = np.arange(32 * nb, dtype=np.uint32)
    dg.vertex_load_sources(src, True)
    st = ctx.stats()
    print("limit %5.1f GB slots %3d batches %4d engine_ms %8.2f  ms/batch/cta %7.2f  fwd %.2f bwd %.2f  -> full run est %7.1f ms" % (
        gb, slots, nb, st["engine_ms"], st["engine_ms"] * slots / nb, st["frac_forward"], st["frac_backward"],
        st["engine_ms"] / nb * 2048), flush=True)
