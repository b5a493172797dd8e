"""One wave of 64-source batches on the G(n,m) N=2^20 graph with 148 / 74 / 37 / 18 resident CTAs: is the
batched engine bound by per-CTA latency (time per wave constant) or by a shared resource (time drops)?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import net_ensembles_b200 as ne
os.environ["NEC_BATCH_WIDTH"] = "64"
g = ne.gnm_scalable(1 << 20, 1 << 22, 7)
rp, col = g.export_csr()
n = len(rp) - 1
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
for slots in (148, 74, 37, 18):
    src = (np.arange(64 * slots, dtype=np.uint64) * 2654435761 % n).astype(np.uint32)
    for rep in range(2):
        dg.vertex_load_sources(src, True)
    st = ctx.stats()
    print("CTAs %3d (slots %d): wave %.1f ms  fwd %.3f bwd %.3f  -> %.2f GTEPS" % (slots, st["n_slots"], st["engine_ms"], st["frac_forward"], st["frac_backward"],
          st["edge_pairs"] / 2 / (st["engine_ms"] * 1e-3) / 1e9), flush=True)
    dg.distance_measures(src)
    print("          forward only (measures): %.1f ms" % ctx.stats()["engine_ms"], flush=True)
