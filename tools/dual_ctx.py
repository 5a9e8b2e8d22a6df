"""Experiment: S contexts on ONE GPU, each with 2^lg / S chains on its own stream, driven from S host threads, against one
context with all the chains.  The kernels of the chain loop are latency bound (issue slots 15-20 % busy), so independent
sub-batches whose phases interleave should overlap.   usage: python tools/dual_ctx.py [lg=20] [steps=32] [S=2]"""
import importlib.util, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(ROOT, "langevin-mcmc_b200", "__init__.py"),
                                              submodule_search_locations=[os.path.join(ROOT, "langevin-mcmc_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["lmc_b200"] = m; spec.loader.exec_module(m)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = m.ParseScene(os.path.join(ROOT, "scenes", os.environ.get("LMC_SCENE", "torus/lmc.xml")))
sc.options["maxdepth"] = int(os.environ.get("LMC_MAXDEPTH", "8"))
chains = 1 << lg
norm, init_small = m.MLTInit(sc, 300000, min(chains, 8192), 32)
init_ls = np.resize(init_small, chains)

def bench(parts):
    per = chains // parts
    ctxs = [m.ChainContext(sc, 0) for _ in range(parts)]
    for g, c in enumerate(ctxs):
        c.begin(per, norm, init_ls, samples_per_chain=steps * 4, chain_base=g * per, total_chains=chains)
    def work(c, n):
        c.run(n); c.synchronize()
    for rep in range(4):
        t = time.time()
        th = [threading.Thread(target=work, args=(c, steps)) for c in ctxs]
        for x in th: x.start()
        for x in th: x.join()
        dt = time.time() - t
        print("parts %d launch %d: %.1f ms  %.2f M mut/s" % (parts, rep, dt * 1e3, chains * steps / dt / 1e6), flush=True)
    for c in ctxs: c.close()

bench(1)
bench(S)
