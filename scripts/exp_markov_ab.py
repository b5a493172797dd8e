"""Same-box A/B (lib/ab_old.so vs lib/ab_new.so) of the Markov-chain shape end to end: 4096 ErC chains
N=256, one m_step + vertex_load(true) of all graphs per step, wall-clock per step through the Python API."""
import sys, os, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import net_ensembles_b200 as ne
    chains = [ne.ErEnsembleC(256, 4.0, 123 + g) for g in range(4096)]
    ctx = ne.Context(0)
    batch = ne.ChainBatch(chains)
    for _ in range(3):
        batch.step_and_load(ctx, 1, True)
    ts, ks = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        batch.step_and_load(ctx, 1, True)
        ts.append((time.perf_counter() - t0) * 1e3)
        ks.append(ctx.stats()["kernel_ms"])
    print("%-7s wall ms/step min %.2f median %.2f   kernel ms %.2f" % (sys.argv[1], min(ts), sorted(ts)[5], min(ks)), flush=True)
else:
    for rep in range(2):
        for tag in ("ab_old", "ab_new"):
            env = dict(os.environ, NEC_LIB_PATH=os.path.join(ROOT, "net_ensembles_b200", "lib", tag + ".so"))
            subprocess.run([sys.executable, __file__, tag], env=env)
