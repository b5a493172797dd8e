"""Wide engine: time of one batch as a function of the team size (CTAs per batch), same batches, on config 4 / 5:
the exponent of the planner's cost model.  python scripts/exp_wide_teams.py erm|ba"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne
which = sys.argv[1] if len(sys.argv) > 1 else "erm"
g = ne.gnm_scalable(1 << 20, 1 << 22, 7) if which == "erm" else ne.ba_scalable(1 << 22, 4, 5, 7)
rp, col = g.export_csr()
n = len(rp) - 1
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
for width in (128, 64):
    os.environ["NEC_WIDE_WIDTH"] = str(width)
    for team in (1, 2, 3, 4, 6, 8):
        os.environ["NEC_WIDE_TEAM"] = str(team)
        nb = 148 // team if which == "erm" else 12          # batches = slots: one batch per slot, whole GPU (erm) / same 12 batches (ba)
        src = (np.arange(nb * width, dtype=np.uint64) * 2654435761 % n).astype(np.uint32)
        ts = []
        for _ in range(2):
            dg.vertex_load_sources(src, True)
            ts.append(ctx.stats()["engine_ms"])
        st = ctx.stats()
        print("%s width %d team %d (asked %d) slots %d batches %d: %.1f ms per wave, %.2f GTEPS" % (
            which, width, st["team_size"], team, st["n_slots"], st["n_batches"], min(ts), st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9), flush=True)
