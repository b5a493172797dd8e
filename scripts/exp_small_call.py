"""Per-call cost of the mid engine on BASELINE configs[0] (ErC N=1000): wall clock per call, kernel time, host issue / wait split.
python scripts/exp_small_call.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne

e = ne.ErEnsembleC(1000, 4.0, 0x5EED)
rp, col = e.graph().export_csr()
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
ctx.set_engine(1)
ref = dg.vertex_load(True)
ctx.set_engine(0)
out = dg.vertex_load(True)
print("engine", ctx.stats()["engine"], "max rel diff vs source-minor engine %.2e" % float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0))))
odd = np.arange(1, 1000, 2, dtype=np.uint32)               # a list that is not a run of consecutive ids
for name, call in (("all sources (consecutive)", lambda: dg.vertex_load(True)), ("500 odd sources", lambda: dg.vertex_load_sources(odd, True))):
    for _ in range(20):
        call()
    t0 = time.perf_counter()
    for _ in range(500):
        call()
    dt = (time.perf_counter() - t0) / 500
    st = ctx.stats()
    print("%s: %.1f us per call, kernel_ms %.4f engine_ms %.4f" % (name, dt * 1e6, st["kernel_ms"], st["engine_ms"]), flush=True)
os.environ["NEC_MID_TIMING"] = "1"
for _ in range(3):
    dg.vertex_load(True)
