"""DRAM traffic of the batched engine's forward pass alone (distance-measure mode) next to the full
vertex_load pass on the G(n,m) N=2^20 graph; run under `ncu --metrics dram__bytes_read.sum,...`."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import net_ensembles_b200 as ne
g = ne.gnm_scalable(1 << 20, 1 << 22, 7)
rp, col = g.export_csr()
n = len(rp) - 1
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
src = (np.arange(9472, dtype=np.uint64) * 2654435761 % n).astype(np.uint32)
dg.distance_measures(src)
st = ctx.stats(); print("measures", st["engine_ms"], st["batch_width"], st["n_slots"], flush=True)
dg.vertex_load_sources(src, True)
st = ctx.stats(); print("load", st["engine_ms"], st["batch_width"], st["n_slots"], st["frac_forward"], flush=True)
