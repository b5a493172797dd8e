"""Timing of the distance-measure mode (diameter / closeness) next to vertex_load."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import net_ensembles_b200 as ne
from tests.helpers import Oracle
from net_ensembles_b200 import _build
ctx = ne.Context(0)
orc = Oracle(_build.ORACLE_PATH)
for name, e in (("Sw N=2^16 p=0.1", ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)), ("ErC N=1000 c=4", ne.ErEnsembleC(1000, 4.0, 123_239_010)),
                ("G(n,m) N=2^20 M=2^22 (4096 sampled sources)", ne.gnm_scalable(1 << 20, 1 << 22, 123_239_010))):
    rp, col = e.export_csr()
    n = len(rp) - 1
    dg = ctx.upload(rp, col)
    src = None if n <= (1 << 16) else np.arange(0, n, n // 4096, dtype=np.uint32)[:4096]
    dg.distance_measures(src)
    t0 = time.perf_counter(); ecc, sumd, reached = dg.distance_measures(src); t1 = time.perf_counter()
    st = ctx.stats()
    k = n if src is None else len(src)
    sample = np.arange(0, k, max(1, k // 64))[:64]
    ssrc = (np.arange(n, dtype=np.uint32) if src is None else src)[sample]
    t2 = time.perf_counter(); recc, rsum, rreach = orc.distance_measures(rp, col, ssrc); t3 = time.perf_counter()
    ok = ecc[sample].tolist() == recc.tolist() and sumd[sample].tolist() == rsum.tolist() and reached[sample].tolist() == rreach.tolist()
    print("%s: %d sources, call %.1f ms (kernel %.1f ms) = %.2f GTEPS; CPU oracle %.2f ms/source (1 thread) -> x%.0f; exact=%s; diameter(sampled max ecc)=%d" % (
        name, k, (t1 - t0) * 1e3, st["kernel_ms"], st["edge_pairs"] / 2 / (t1 - t0) / 1e9, (t3 - t2) * 1e3 / len(ssrc),
        (t3 - t2) / len(ssrc) * k / (t1 - t0), ok, int(ecc.max())), flush=True)
    dg.free()
