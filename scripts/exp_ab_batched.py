"""Same-box A/B of two library builds (lib/ab_old.so, lib/ab_new.so) on the batched engine:
G(n,m) N=2^20 (9472 sampled sources) and, with `ba` as first argument, Barabasi-Albert N=2^22 (3072)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 2:
    import numpy as np
    import net_ensembles_b200 as ne
    which, tag = sys.argv[1], sys.argv[2]
    g = ne.gnm_scalable(1 << 20, 1 << 22, 7) if which == "erm" else ne.ba_scalable(1 << 22, 4, 5, 7)
    rp, col = g.export_csr()
    n = len(rp) - 1
    ctx = ne.Context(0)
    dg = ctx.upload(rp, col)
    ns = int(os.environ.get("AB_NS", 9472 if which == "erm" else 3072))
    src = (np.arange(ns, dtype=np.uint64) * 2654435761 % n).astype(np.uint32)
    ts = []
    for _ in range(3):
        dg.vertex_load_sources(src, True)
        ts.append(ctx.stats()["engine_ms"])
    st = ctx.stats()
    print("%-4s %-7s engine_ms %s  width %d slots %d fwd %.3f bwd %.3f  GTEPS %.2f" % (
        which, tag, ["%.1f" % t for t in ts], st.get("batch_width", 0), st["n_slots"], st["frac_forward"], st["frac_backward"],
        st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9), flush=True)
else:
    which = sys.argv[1] if len(sys.argv) > 1 else "erm"
    for rep in range(int(os.environ.get("AB_REPS", "2"))):
        for tag in os.environ.get("AB_TAGS", "ab_old,ab_new").split(","):
            env = dict(os.environ, NEC_LIB_PATH=os.path.join(ROOT, "net_ensembles_b200", "lib", tag + ".so"))
            subprocess.run([sys.executable, __file__, which, tag], env=env)
