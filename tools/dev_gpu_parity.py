import sys, time, ctypes, importlib.util, numpy as np
spec = importlib.util.spec_from_file_location("lmc_b200", "/root/repo/langevin-mcmc_b200/__init__.py", submodule_search_locations=["/root/repo/langevin-mcmc_b200"])
m = importlib.util.module_from_spec(spec); sys.modules["lmc_b200"] = m; spec.loader.exec_module(m)
sc = m.ParseScene("/root/repo/scenes/torus/lmc.xml")
print(sc.info)
sc.options["maxdepth"] = 4
nch, steps = 1024, 100
norm, initls = m.MLTInit(sc, 300000, nch, 32)
print("norm", norm)
ctx = m.ChainContext(sc, 0)
ctx.begin(nch, norm, initls, samples_per_chain=steps)
t = time.time(); tr, a = ctx.run(steps, trace=True, a_trace=True); dt = time.time() - t
st = ctx.stats(); print("gpu", dt, st)
film = ctx.film(); print("film sum", film.sum())
# oracle
L = ctypes.CDLL('/root/repo/oracle/liblmc_oracle.so'); L.lmco_scene_load.restype = ctypes.c_void_p
h = ctypes.c_void_p(L.lmco_scene_load(b'/root/repo/scenes/torus/lmc.xml'))
L.lmco_set_option(h, b'maxdepth', ctypes.c_double(4))
vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)
ofilm = np.zeros_like(film); otr = np.zeros_like(tr); oa = np.zeros_like(a); ostats = np.zeros(10, np.uint64)
t = time.time()
L.lmco_run_chains(h, nch, 0, nch, ctypes.c_longlong(steps), ctypes.c_longlong(steps), ctypes.c_float(norm), vp(initls), vp(ofilm), vp(otr), vp(oa), 8, vp(ostats))
print("cpu", time.time() - t, ostats)
print("trace equal:", np.array_equal(tr, otr), "mismatching chains:", int((tr != otr).any(axis=1).sum()), "a bit-equal:", np.array_equal(a.view(np.uint32), oa.view(np.uint32)))
print("film sums", film.sum(), ofilm.sum(), "max abs diff", np.abs(film - ofilm).max())
print("grad evals gpu", st["gradient_evals"], "cpu", ostats[8], ostats[9])
# throughput probe at depth 8
sc2 = m.ParseScene("/root/repo/scenes/torus/lmc.xml")
for nch2, steps2 in ((1 << 14, 64), (1 << 17, 64), (1 << 20, 32)):
    norm2, init2 = m.MLTInit(sc2, max(300000, 4 * nch2), nch2, 32) if nch2 <= (1 << 17) else (norm2, np.resize(init2, nch2))
    c2 = m.ChainContext(sc2, 0)
    c2.begin(nch2, norm2, init2, samples_per_chain=steps2 * 4)
    c2.run(steps2); c2.synchronize()
    t = time.time(); c2.run(steps2); c2.synchronize(); dt = time.time() - t
    s2 = c2.stats()
    print("chains %d steps %d: %.3f s  %.2f M mut/s  kernel ms %.1f  accept %s grad %d" % (nch2, steps2, dt, nch2 * steps2 / dt / 1e6, s2["last_kernel_ms"], s2["accepted"], s2["gradient_evals"]))
    c2.close()
