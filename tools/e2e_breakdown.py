"""Where the end-to-end job time goes: begin (H2D + chain init), first iteration (all large steps),
steady iterations, film read.  usage: python tools/e2e_breakdown.py [chains_log2=20]"""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(ROOT, "langevin-mcmc_b200", "__init__.py"),
                                              submodule_search_locations=[os.path.join(ROOT, "langevin-mcmc_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["lmc_b200"] = m; spec.loader.exec_module(m)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sc = m.ParseScene(os.path.join(ROOT, "scenes", "torus", "lmc.xml"))
sc.options["maxdepth"] = 8
chains = 1 << lg
ctx = m.ChainContext(sc, 0)
norm, init_ls = ctx.mlt_init(max(300000, 4 * chains), chains, 65536)
def t(f):
    ctx.synchronize(); t0 = time.time(); r = f(); ctx.synchronize(); return (time.time() - t0) * 1e3, r
for rep in range(3):
    tb, _ = t(lambda: ctx.begin(chains, norm, init_ls, samples_per_chain=32))
    t1, _ = t(lambda: ctx.run(1))
    t2, _ = t(lambda: ctx.run(1))
    t3, _ = t(lambda: ctx.run(30))
    tf, _ = t(lambda: ctx.film())
    print("rep %d: begin %.1f ms, iteration 0 %.1f ms, iteration 1 %.1f ms, 30 iterations %.1f ms (%.2f each), film read %.1f ms" % (rep, tb, t1, t2, t3, t3 / 30, tf))
