"""Wide (compact batched) engine against the mid engine on BASELINE configs[1] (small-world N=2^16, all sources):
same sources, same box.  python scripts/exp_wide_sw.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import net_ensembles_b200 as ne

e = ne.SwEnsemble(1 << 16, 0.1, 123_239_010, 2)
rp, col = e.export_csr()
n = len(rp) - 1
ctx = ne.Context(0)
dg = ctx.upload(rp, col)
ref = None


def run(tag, engine, env):
    global ref
    for k in ("NEC_WIDE_WIDTH", "NEC_WIDE_TEAM", "NEC_WIDE_QUEUE", "NEC_WIDE_THREADS", "NEC_PULL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx.set_engine(engine)
    ts = []
    try:
        for _ in range(3):
            out = dg.vertex_load(True)
            st = ctx.stats()
            ts.append(st["kernel_ms"])
    except Exception as ex:                       # noqa: BLE001
        print("%s: FAILED %s" % (tag, ex), flush=True)
        return
    if ref is None:
        ref = out
    err = float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0)))
    print("%s: kernel_ms %s  GTEPS %.1f  engine %s width %d team %d slots %d requeued %d depth %d pull/push %d/%d phases %.3f/%.3f/%.3f  max rel diff vs mid %.2e" % (
        tag, ["%.2f" % t for t in ts], st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9, st["engine"], st["batch_width"], st["team_size"], st["n_slots"],
        st.get("requeued_batches", -1), st["max_depth"], st.get("pull_levels", 0), st.get("push_levels", 0),
        st["frac_reset"], st["frac_forward"], st["frac_backward"], err), flush=True)


run("mid", 2, {})
for q in ("12", "24", "48"):
    run("wide q=%s" % q, 3, {"NEC_WIDE_QUEUE": q})
for team in ("1", "2", "4", "8"):
    run("wide q=48 team=%s" % team, 3, {"NEC_WIDE_QUEUE": "48", "NEC_WIDE_TEAM": team})
for w in ("128", "64"):
    run("wide q=48 width=%s" % w, 3, {"NEC_WIDE_QUEUE": "48", "NEC_WIDE_WIDTH": w})
for pm in ("0", "2"):
    run("wide q=48 pull=%s" % pm, 3, {"NEC_WIDE_QUEUE": "48", "NEC_PULL": pm})
