"""BASELINE configs[3] in full on one GPU: vertex_load(true) of the G(n,m) N=2^20 M=2^22 graph over ALL 2^20
sources in one call (no sampling); prints the engine time, GTEPS and an oracle spot check."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import net_ensembles_b200 as ne
g = ne.gnm_scalable(1 << 20, 1 << 22, 123_239_010)
rp, col = g.export_csr()
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
t0 = time.perf_counter()
load = dg.vertex_load(True)
wall = time.perf_counter() - t0
st = ctx.stats()
te = st["edge_pairs"] / 2
print("all %d sources: wall %.2f s, engine %.2f s, %.2f GTEPS (engine), %.2f GTEPS (call); slots %d width %d batches %d depth %d" % (
    st["n_sources"], wall, st["engine_ms"] * 1e-3, te / (st["engine_ms"] * 1e-3) / 1e9, te / wall / 1e9,
    st["n_slots"], st["batch_width"], st["n_batches"], st["max_depth"]), flush=True)
b_alg = 24.0 * st["edge_pairs"] + 40.0 * st["reached_pairs"]
print("algorithmic bytes %.3e -> %.1f GB/s = %.3f of 6547.2 GB/s" % (b_alg, b_alg / (st["engine_ms"] * 1e-3) / 1e9, b_alg / (st["engine_ms"] * 1e-3) / 1e9 / 6547.2))
print("sum of loads %.17g  (= sum of all BFS distances)" % float(load.sum()), " min %.6g max %.6g" % (load.min(), load.max()))
