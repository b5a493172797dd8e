"""Steady-state cost of the structural probe (engine_hint): upload the same graph again and again, ask for the plan.
python scripts/exp_probe_cost.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne
ctx = ne.Context(0)
for name, g in (("small-world N=2^16", ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)), ("G(n,m) N=2^16 degree 8", ne.gnm_scalable(1 << 16, 1 << 18, 3)),
                ("small-world N=150000", ne.SwEnsemble(150000, 0.1, 5, 2))):
    rp, col = g.export_csr()
    for mode in ("probe", "size rule only (NEC_PROBE=0)"):
        if mode != "probe":
            os.environ["NEC_PROBE"] = "0"
        ts = []
        for _ in range(6):
            dg = ctx.upload(rp, col)
            t0 = time.perf_counter()
            w = dg.wave_size(1 << 16)
            ts.append((time.perf_counter() - t0) * 1e3)
            dg.free()
        os.environ.pop("NEC_PROBE", None)
        print("%s, %s: wave %d, plan call took %s ms" % (name, mode, w, ["%.2f" % t for t in ts]), flush=True)
