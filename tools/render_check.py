"""Full render (direct pre-pass + LMC chains + MergeBuffer) of a bundled scene at a given spp and its distance
to the render the reference ships (tests/golden/reference_images.npz).  usage: render_check.py torus|door spp"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_package, SCENES, GOLDEN
os.environ.pop("LMC_WAVEFRONT", None)      # conftest forces the wavefront form for the parity tests; a render uses the library's own choice
m = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "torus"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 245
xml = os.path.join(SCENES, "torus" if name == "torus" else "veachdoor", "lmc.xml")
sc = m.ParseScene(xml)
W, H = sc.width, sc.height
ctx = m.ChainContext(sc, 0)
t0 = time.time()
dspp = 64
direct = ctx.direct_lighting(dspp)
t1 = time.time()
chains = 1 << 16
steps = int(np.ceil(spp * W * H / chains))
norm, init_ls = ctx.mlt_init(max(300000, 4 * chains), chains, 65536)
ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
ctx.run(steps); ctx.synchronize()
t2 = time.time()
film = m.MergeBuffer(direct, 1.0 / dspp, ctx.film(), 1.0 / (chains * steps / float(W * H)))
ref = np.load(os.path.join(GOLDEN, "reference_images.npz"))["torus_lmc" if name == "torus" else "door_lmc"]
h, w, _ = film.shape
img = film[: h // 8 * 8, : w // 8 * 8].reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))
print("%s %d spp: direct pass %.2f s (%d spp), chains %.2f s (%d chains x %d mutations = %.1f M mut/s), mean ratio %.4f, relMSE(8x8 box) %.5f" % (
    name, spp, t1 - t0, dspp, t2 - t1, chains, steps, chains * steps / (t2 - t1) / 1e6, img.mean() / ref.mean(),
    float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-2)))))
