"""Same-box A/B of two library builds on the headline workload (mid engine)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np
    import net_ensembles_b200 as ne
    e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
    rp, col = e.export_csr()
    ctx = ne.Context(0)
    dg = ctx.upload(rp, col)
    src = np.arange(0, 1 << 16, 4, dtype=np.uint32)
    dg.vertex_load_sources(src[:2000], True)
    ts = []
    for _ in range(3):
        dg.vertex_load_sources(src, True)
        ts.append(ctx.stats()["engine_ms"])
    st = ctx.stats()
    print("%-8s engine_ms %s -> full-run est %.1f ms  fwd %.3f bwd %.3f" % (sys.argv[1], ["%.2f" % t for t in ts], min(ts) * 4, st["frac_forward"], st["frac_backward"]), flush=True)
else:
    for rep in range(2):
        for tag in ("ab_old", "ab_new"):
            env = dict(os.environ, NEC_LIB_PATH=os.path.join(ROOT, "net_ensembles_b200", "lib", tag + ".so"))
            subprocess.run([sys.executable, __file__, tag], env=env)
