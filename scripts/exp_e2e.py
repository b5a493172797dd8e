"""Where does the host-API (e2e) time go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import net_ensembles_b200 as ne
e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
rp, col = e.export_csr()
ctx = ne.Context(0)
src = np.arange(1 << 16, dtype=np.uint32)
for it in range(3):
    t0 = time.perf_counter(); dg = ctx.upload(rp, col); t1 = time.perf_counter()
    out = dg.vertex_load_sources(src, True); t2 = time.perf_counter()
    st = ctx.stats()
    dg.free(); t3 = time.perf_counter()
    print("upload %.1f ms  load call %.1f ms (kernel %.1f ms)  free %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st["kernel_ms"], (t3 - t2) * 1e3), flush=True)
t0 = time.perf_counter(); out = e.vertex_load(True, ctx); t1 = time.perf_counter()
print("ensemble.vertex_load(true) %.1f ms" % ((t1 - t0) * 1e3))
