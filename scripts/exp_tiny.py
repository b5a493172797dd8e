"""How long does one warp of small_graph_kernel take per source on a single graph (1 CTA, 8 warps, nothing else on the GPU),
against the mid engine's whole call on the same graph?  python scripts/exp_tiny.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne
ctx = ne.Context(0)
for n in (128, 256, 500):
    e = ne.ErEnsembleC(n, 4.0, 0x5EED + n)
    rp, col = e.graph().export_csr()
    for _ in range(3):
        outs = ne.vertex_load_batch(ctx, [(rp, col)], True)
    st = ctx.stats()
    t_small = st["kernel_ms"]
    dg = ctx.upload(rp, col)
    for _ in range(3):
        ref = dg.vertex_load(True)
    st2 = ctx.stats()
    print("n %d: small kernel, one CTA: %.1f us for %d sources = %.2f us per source per warp (8 warps); mid engine call: engine %.1f us kernel %.1f us; max rel diff %.1e" % (
        n, 1e3 * t_small, n, 1e3 * t_small / (n / 8), 1e3 * st2["engine_ms"], 1e3 * st2["kernel_ms"],
        float(np.max(np.abs(outs[0] - ref) / np.maximum(np.abs(ref), 1.0)))), flush=True)
    dg.free()
