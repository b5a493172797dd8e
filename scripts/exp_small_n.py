"""Which traversal engine should serve graphs between the one-CTA-per-graph kernel (n <= 512) and the mid
engine's default range (n >= 2048)?  ErC c=4 at several n, all sources, engine 1 (batched) vs 2 (mid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import net_ensembles_b200 as ne
ctx = ne.Context(0)
for n in (300, 600, 1000, 1500, 2048, 4096):
    e = ne.ErEnsembleC(n, 4.0, 11)
    rp, col = e.export_csr()
    dg = ctx.upload(rp, col)
    row = []
    for eng in (1, 2):
        ctx.set_engine(eng)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); dg.vertex_load(True); best = min(best, time.perf_counter() - t0)
        st = ctx.stats()
        row.append("engine %d(ran %d): call %.3f ms kernel %.3f ms" % (eng, st["engine"] & 15, best * 1e3, st["kernel_ms"]))
    ctx.set_engine(0)
    dg.free()
    print("n=%5d  %s" % (n, "   ".join(row)), flush=True)
