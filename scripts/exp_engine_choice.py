"""Mid engine (one source per CTA) against the compact batched engine on graphs below the size threshold: where does the
crossover lie for random graphs, as opposed to the small-world config the threshold was set on?  python scripts/exp_engine_choice.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne

ctx = ne.Context(0)
cases = [("gnm n=2^13 deg 8", lambda: ne.gnm_scalable(1 << 13, 1 << 15, 3)), ("gnm n=2^14 deg 8", lambda: ne.gnm_scalable(1 << 14, 1 << 16, 3)),
         ("gnm n=2^15 deg 8", lambda: ne.gnm_scalable(1 << 15, 1 << 17, 3)), ("gnm n=2^16 deg 8", lambda: ne.gnm_scalable(1 << 16, 1 << 18, 3)),
         ("gnm n=2^17 deg 8", lambda: ne.gnm_scalable(1 << 17, 1 << 19, 3)), ("gnm n=2^16 deg 4", lambda: ne.gnm_scalable(1 << 16, 1 << 17, 3)),
         ("gnm n=2^16 deg 16", lambda: ne.gnm_scalable(1 << 16, 1 << 19, 3)), ("ba n=2^16 m=4", lambda: ne.ba_scalable(1 << 16, 4, 5, 3)),
         ("sw n=2^14 p=0.1", lambda: ne.SwEnsemble(1 << 14, 0.1, 7, 2)), ("sw n=2^16 p=0.1", lambda: ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)),
         ("sw n=2^16 p=0.3", lambda: ne.SwEnsemble(1 << 16, 0.3, 5, 2)), ("sw n=2^16 p=0.6", lambda: ne.SwEnsemble(1 << 16, 0.6, 5, 2)),
         ("sw n=2^16 p=0.02", lambda: ne.SwEnsemble(1 << 16, 0.02, 5, 2)), ("sw n=2^12 p=0.1", lambda: ne.SwEnsemble(1 << 12, 0.1, 5, 2)),
         ("gnm n=2^16 deg 3", lambda: ne.gnm_scalable(1 << 16, 3 << 15, 3)), ("gnm n=2^12 deg 8", lambda: ne.gnm_scalable(1 << 12, 1 << 14, 3)),
         ("gnm n=2^12 deg 4", lambda: ne.gnm_scalable(1 << 12, 1 << 13, 3)), ("ba n=2^13 m=2", lambda: ne.ba_scalable(1 << 13, 2, 3, 3)),
         ("erc n=20000 c=2.5", lambda: ne.ErEnsembleC(20000, 2.5, 5)), ("sw n=150000 p=0.1", lambda: ne.SwEnsemble(150000, 0.1, 5, 2)),
         ("gnm n=150000 deg 6", lambda: ne.gnm_scalable(150000, 450000, 3))]
os.environ["NEC_WIDE_QUEUE"] = "64"
for name, make in cases:
    g = make()
    rp, col = g.export_csr()
    n = len(rp) - 1
    dg = ctx.upload(rp, col)
    ns = min(n, 37 * 256 * 2)
    src = (np.arange(ns, dtype=np.uint64) * 2654435761 % n).astype(np.uint32) if ns < n else np.arange(n, dtype=np.uint32)
    res = {}
    ctx.set_engine(3)                                   # the probe a planner could run: one batch, vertex ids 0..255
    dg.vertex_load_sources(np.arange(min(n, 256), dtype=np.uint32), True)
    pst = ctx.stats()
    probe_e, probe_ms = pst["vertex_visits"] / max(1, pst["n_batches"] * n), pst["kernel_ms"]
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        src = np.arange(n, dtype=np.uint32)
        ns = n
    for eng, tag in ((2, "mid"), (3, "wide")):
        ctx.set_engine(eng)
        ts = []
        for _ in range(3):
            out = dg.vertex_load_sources(src, True)
            ts.append(ctx.stats()["kernel_ms"])
        st = ctx.stats()
        res[tag] = (min(ts), st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9, out, st)
    ctx.set_engine(0)
    err = float(np.max(np.abs(res["mid"][2] - res["wide"][2]) / np.maximum(np.abs(res["mid"][2]), 1.0)))
    sw = res["wide"][3]
    print("%-18s %6d sources: mid %8.2f ms %6.1f GTEPS | wide %8.2f ms %6.1f GTEPS (width %d team %d slots %d requeued %d depth %d entries/vertex %.1f) | wide/mid speed %.2f  max rel diff %.1e | probe: entries/vertex %.1f, %.2f ms" % (
        name, ns, res["mid"][0], res["mid"][1], res["wide"][0], res["wide"][1], sw["batch_width"], sw["team_size"], sw["n_slots"], sw["requeued_batches"],
        sw["max_depth"], sw["vertex_visits"] / max(1, sw["n_batches"] * n), res["mid"][0] / res["wide"][0], err, probe_e, probe_ms), flush=True)
    dg.free()
