"""Does the L1 size (= 256 KB minus the shared-memory carve-out) matter to the mid engine?  Small-world N=2^16, a quarter of the
sources, preferred carve-out forced through NEC_MID_CARVEOUT (percent of the maximum).  python scripts/exp_carveout.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne

e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
rp, col = e.export_csr()
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
src = np.arange(0, 1 << 16, 4, dtype=np.uint32)
ref = None
for pct in ("default", "100", "86", "75", "72", "default"):
    if pct == "default":
        os.environ.pop("NEC_MID_CARVEOUT", None)
    else:
        os.environ["NEC_MID_CARVEOUT"] = pct
    ts = []
    for _ in range(3):
        out = dg.vertex_load_sources(src, True)
        ts.append(ctx.stats()["engine_ms"])
    st = ctx.stats()
    if ref is None:
        ref = out
    print("mid carve-out %s: engine_ms %s slots %d phases %.3f/%.3f/%.3f %s" % (pct, ["%.2f" % t for t in ts], st["n_slots"], st["frac_reset"], st["frac_forward"], st["frac_backward"],
          "bitwise-equal" if out.tobytes() == ref.tobytes() else "DIFFERS"), flush=True)
