"""BASELINE config 3: 4096 ErC chains N=256 c=4, one m_step each, vertex_load(true) of all graphs in one launch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import net_ensembles_b200 as ne
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = time.perf_counter()
chains = [ne.ErEnsembleC(256, 4.0, 123_239_010 + g) for g in range(G)]
print("built %d chains in %.1f s" % (G, time.perf_counter() - t0), flush=True)
ctx = ne.Context(0)
for step in range(3):
    t0 = time.perf_counter()
    for c in chains:
        c.m_step()
    csr = [c.export_csr() for c in chains]
    graphs = [(rp.astype(np.uint32), col) for rp, col in csr]
    t1 = time.perf_counter()
    outs = ne.vertex_load_batch(ctx, graphs, True)
    t2 = time.perf_counter()
    st = ctx.stats()
    te = st["edge_pairs"] / 2
    print("step %d: host m_step+export %.1f ms, batch call %.1f ms (kernel %.2f ms) -> %.2f GTEPS kernel, %.2f GTEPS call; depth %d" % (
        step, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["kernel_ms"], te / st["kernel_ms"] / 1e6, te / (t2 - t1) / 1e9, st["max_depth"]), flush=True)

for step in range(3):
    t0 = time.perf_counter()
    outs = ne.chains_step_and_load(ctx, chains, 1, True)
    t1 = time.perf_counter()
    st = ctx.stats()
    print("C++ path step %d: m_step + export + load %.1f ms (kernel %.2f ms) -> %.2f GTEPS end to end" % (
        step, (t1 - t0) * 1e3, st["kernel_ms"], st["edge_pairs"] / 2 / (t1 - t0) / 1e9), flush=True)
