"""Soak run: 2^20 chains x N iterations in launches of 64 (default N = 1024, ~1e9 mutations): counters must add up,
the film must stay finite, throughput per launch is printed (sustained clocks / power).  usage: soak.py [N=1024]"""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(ROOT, "langevin-mcmc_b200", "__init__.py"),
                                              submodule_search_locations=[os.path.join(ROOT, "langevin-mcmc_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["lmc_b200"] = m; spec.loader.exec_module(m)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sc = m.ParseScene(os.path.join(ROOT, "scenes", "torus", "lmc.xml")); sc.options["maxdepth"] = 8
chains = 1 << 20
ctx = m.ChainContext(sc, 0)
norm, init_ls = ctx.mlt_init(4 * chains, chains, 65536)
ctx.begin(chains, norm, init_ls, samples_per_chain=N)
rates = []
for k in range(N // 64):
    t = time.time(); ctx.run(64); ctx.synchronize(); dt = time.time() - t
    rates.append(chains * 64 / dt / 1e6)
st = ctx.stats()
film = ctx.film()
assert sum(st["proposed"]) == chains * (N // 64) * 64, st
assert np.isfinite(film).all()
mean = float(film.sum()) / (chains * (N // 64) * 64) / 3.0
print("soak ok: %.2e mutations, M mut/s per launch: first %.1f min %.1f median %.1f last %.1f; accept rates L/S/M = %.3f %.3f %.3f; "
      "gradient evals %d (non-finite %d); film mean / normalization = %.4f" % (
          float(sum(st["proposed"])), rates[0], min(rates), float(np.median(rates)), rates[-1],
          st["accepted"][0] / max(1, st["proposed"][0]), st["accepted"][1] / max(1, st["proposed"][1]),
          st["accepted"][3] / max(1, st["proposed"][3]), st["gradient_evals"], st["gradient_nonfinite"], mean / norm))
