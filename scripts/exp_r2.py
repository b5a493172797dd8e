"""Round-2 same-box A/B runs of the batched engine: one process per graph, the graph uploaded once, then every
variant (a set of environment knobs the library reads per call, or another build via a child process) timed on
the same sampled sources.

  python scripts/exp_r2.py erm|ba [tag:ENV=V,ENV=V ...]

Default variants compare the forward-pass direction: NEC_PULL=0 (push only), 1 (cost model), 2 (pull only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import net_ensembles_b200 as ne  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "erm"
    if len(sys.argv) > 3 and sys.argv[2] == "--lib":          # another build of the library (lib/<tag>.so): child process
        import subprocess
        env = dict(os.environ, NEC_LIB_PATH=os.path.join(ROOT, "net_ensembles_b200", "lib", sys.argv[3] + ".so"), AB_LIB_TAG=sys.argv[3])
        raise SystemExit(subprocess.run([sys.executable, __file__, which] + sys.argv[4:], env=env).returncode)
    specs = sys.argv[2:] or ["push:NEC_PULL=0", "auto:NEC_PULL=1", "pull:NEC_PULL=2"]
    g = ne.gnm_scalable(1 << 20, 1 << 22, 7) if which == "erm" else ne.ba_scalable(1 << 22, 4, 5, 7)
    rp, col = g.export_csr()
    n = len(rp) - 1
    ctx = ne.Context(0)
    dg = ctx.upload(rp, col)
    waves = int(os.environ.get("AB_WAVES", "1"))
    ns = int(os.environ.get("AB_NS", waves * dg.wave_size(n)))
    src = (np.arange(ns, dtype=np.uint64) * 2654435761 % n).astype(np.uint32)
    ref = None
    for spec in specs:
        tag, _, envs = spec.partition(":")
        pairs = [kv.split("=", 1) for kv in envs.split(",") if kv]
        for k, v in pairs:
            os.environ[k] = v
        ts = []
        for _ in range(int(os.environ.get("AB_REPS", "3"))):
            out = dg.vertex_load_sources(src, True)
            ts.append(ctx.stats()["engine_ms"])
        st = ctx.stats()
        same = "first" if ref is None else ("bitwise-equal" if out.tobytes() == ref.tobytes() else "DIFFERS max rel %.3g" % float(np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1.0))))
        if ref is None:
            ref = out.copy()
        print(os.environ.get("AB_LIB_TAG", "product") + " %-4s %-10s ns %d engine_ms %s width %d slots %d team %s fwd %.3f bwd %.3f GTEPS %.2f frac %.3f  %s sha %s" % (
            which, tag, ns, ["%.1f" % t for t in ts], st.get("batch_width", 0), st["n_slots"], st.get("team_size", "?"),
            st["frac_forward"], st["frac_backward"], st["edge_pairs"] / 2 / (min(ts) * 1e-3) / 1e9,
            (24.0 * st["edge_pairs"] + 40.0 * st["reached_pairs"]) / (min(ts) * 1e-3) / 1e9 / 6547.2, same, __import__('hashlib').sha1(out.tobytes()).hexdigest()[:10]), flush=True)
        for k, _ in pairs:
            os.environ.pop(k, None)
    dg.free()
    ctx.close()


if __name__ == "__main__":
    main()
