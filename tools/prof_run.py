"""Small fixed workload for ncu captures: torus, LMC, maxdepth 8.
usage: python tools/prof_run.py [chains_log2=16] [steps=8] [launches=3]"""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(ROOT, "langevin-mcmc_b200", "__init__.py"),
                                              submodule_search_locations=[os.path.join(ROOT, "langevin-mcmc_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["lmc_b200"] = m; spec.loader.exec_module(m)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = m.ParseScene(os.path.join(ROOT, "scenes", os.environ.get("LMC_SCENE", "torus/lmc.xml")))
sc.options["maxdepth"] = int(os.environ.get("LMC_MAXDEPTH", "8"))
for kv in filter(None, os.environ.get("LMC_OPTS", "").split(",")):      # e.g. LMC_OPTS=globalcache=1,adjointcompat=0
    k, v = kv.split("=")
    sc.options[k] = float(v)
chains = 1 << lg
norm, init_small = m.MLTInit(sc, 300000, min(chains, 8192), 32)
init_ls = np.resize(init_small, chains)
ctx = m.ChainContext(sc, 0)
ctx.begin(chains, norm, init_ls, samples_per_chain=steps * launches)
for k in range(launches):
    t = time.time(); ctx.run(steps); ctx.synchronize(); dt = time.time() - t
    print("launch %d: %.1f ms  %.2f M mut/s" % (k, dt * 1e3, chains * steps / dt / 1e6))
print(ctx.stats())
