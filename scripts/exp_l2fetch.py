"""A/B: cudaLimitMaxL2FetchGranularity (NEC_L2_FETCH = 32 / 64 / 128) on the mid engine (small-world N=2^16)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne
e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
rp, col = e.export_csr()
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
src = np.arange(0, 1 << 16, 2, dtype=np.uint32)
dg.vertex_load_sources(src[:2000], True)
ref = None
for gran in ("64", "32", "128", "64", "32"):
    os.environ["NEC_L2_FETCH"] = gran
    ts = []
    for _ in range(3):
        out = dg.vertex_load_sources(src, True)
        ts.append(ctx.stats()["engine_ms"])
    if ref is None:
        ref = out.copy()
    print("mid sw65536 half the sources, L2 fetch %s B: engine_ms %s  %s" % (gran, ["%.2f" % t for t in ts], "bitwise-equal" if out.tobytes() == ref.tobytes() else "DIFFERS"), flush=True)
