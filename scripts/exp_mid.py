"""Experiment: mid engine time on the small-world workload for 1/2/3 CTAs per SM."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    import net_ensembles_b200 as ne
    e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
    rp, col = e.export_csr()
    ctx = ne.Context(0)
    dg = ctx.upload(rp, col)
    src = np.arange(0, 1 << 16, 4, dtype=np.uint32)        # 16384 sources
    dg.vertex_load_sources(src[:2000], True)
    dg.vertex_load_sources(src, True)
    st = ctx.stats()
    print("ctas/sm %s slots %d engine_ms %.2f -> full-run est %.1f ms  phases r/f/b %.3f %.3f %.3f" % (
        sys.argv[1], st["n_slots"], st["engine_ms"], st["engine_ms"] * 4, st["frac_reset"], st["frac_forward"], st["frac_backward"]), flush=True)
else:
    for k in ("1", "2", "3"):
        env = dict(os.environ, NEC_MID_CTAS_PER_SM=k)
        subprocess.run([sys.executable, __file__, k], env=env)
