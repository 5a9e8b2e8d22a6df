"""Development aid: finds chains whose splats make a film value negative in a 2^20-chain H2MC run and replays
them on the CPU oracle (result so far: the oracle produces the same negative value bit for bit -- reference
arithmetic, not a device bug)."""
import importlib.util, os, sys, time
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Oracle, load_package
m = load_package()
xml = os.path.join(ROOT, "scenes", "torus", "h2mc.xml")
sc = m.ParseScene(xml); sc.options["maxdepth"] = 8
chains = 1 << 20; steps = 3
ctx = m.ChainContext(sc, 0)
norm, init_small = ctx.mlt_init(300000, 4096, 4096)
init_ls = np.resize(init_small, chains)
B = 1 << 16
o = Oracle(); h = o.load(xml); o.set_option(h, "maxdepth", 8)
found = 0
for b in range(chains // B):
    ctx.begin(B, norm, init_ls, chain_base=b * B, total_chains=chains, samples_per_chain=steps)
    ctx.run(steps)
    film = ctx.film()
    if film.min() < 0:
        idx = np.unravel_index(np.argmin(film), film.shape)
        print("block", b, "negative", film.min(), "at", idx, flush=True)
        # narrow down to 1024 chains
        for s in range(B // 1024):
            ctx.begin(1024, norm, init_ls, chain_base=b * B + s * 1024, total_chains=chains, samples_per_chain=steps)
            tr, a = ctx.run(steps, trace=True, a_trace=True)
            f2 = ctx.film()
            if f2.min() < 0:
                of, otr, oa, ost = o.run_chains(h, 1024, steps, norm, init_ls, chain_base=b * B + s * 1024, total_chains=chains, samples_per_chain=steps, threads=16)
                print("  sub", s, "gpu min", f2.min(), "oracle min", of.min(), "trace equal", np.array_equal(tr, otr), "a equal", np.array_equal(a.view(np.uint32), oa.view(np.uint32)),
                      "film maxdiff", np.abs(f2 - of).max(), flush=True)
                found += 1
        if found >= 2: break
print("done, found", found)
