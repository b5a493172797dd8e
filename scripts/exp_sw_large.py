"""Small-world graph beyond the mid engine's size through the default plan: how many batches does the compact engine hand back?
python scripts/exp_sw_large.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne
ctx = ne.Context(0)
g = ne.SwEnsemble(1 << 18, 0.1, 11, 2)
rp, col = g.export_csr()
n = len(rp) - 1
dg = ctx.upload(rp, col)
src = np.arange(0, 37 * 256 * 2, dtype=np.uint32)             # consecutive ids, as an all-source job batches them
for tag, env in (("default plan", {}), ("queue of 12 (the old default)", {"NEC_WIDE_QUEUE": "12"}), ("queue of 48", {"NEC_WIDE_QUEUE": "48"})):
    for k in ("NEC_WIDE_QUEUE",):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for _ in range(2):
        out = dg.vertex_load_sources(src, True)
        ts.append(ctx.stats()["kernel_ms"])
    st = ctx.stats()
    print("%s: kernel_ms %s GTEPS %.1f engine %d width %d team %d slots %d requeued %d of %d batches, entries/vertex %.1f, workspace %.1f GB" % (
        tag, ["%.1f" % t for t in ts], st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9, st["engine"], st["batch_width"], st["team_size"], st["n_slots"],
        st["requeued_batches"], st["n_batches"], st["vertex_visits"] / max(1, st["n_batches"] * n), st["workspace_bytes"] / 1e9), flush=True)
