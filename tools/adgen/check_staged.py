"""Composition of the per-vertex stage programs == the whole-path program (both sweeps in the reference's order)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import check_whole as cw
import stages as st

ROOT = cw.ROOT
g = np.load(os.path.join(ROOT, "tests", "golden", "path_golden.npz"))
step = int(sys.argv[1]) if len(sys.argv) > 1 else 9
progs = {}
worst = 0.0
for i in range(0, len(g["c"]), step):
    c, l = int(g["c"][i]), int(g["l"][i])
    if (c, l) not in progs:
        progs[(c, l)] = cw.build(c, l)
    prog, dim = progs[(c, l)]
    vals = {}
    for name in prog.inputs:
        arr, idx = name[:-1].split("[")
        src = {"primary": g["primary"][i], "scene": g["scene_ser"][i], "vert": g["vert"][i]}[arr]
        vals[name] = float(src[int(idx)])
    for compat in (True, False):
        outs, adj = prog.evaluate(vals, [1.0], compat=compat)
        gw = np.array([adj["primary[%d]" % (k + 1)] for k in range(dim)])
        v, gs = st.path_grad_staged(c, l, g["scene_ser"][i], g["primary"][i], g["vert"][i], compat)
        gs = np.array(gs)
        if not np.isfinite(gw).all():
            continue
        e = np.linalg.norm(gs - gw) / (np.linalg.norm(gw) + 1e-30)
        worst = max(worst, e)
        if e > 1e-9 or abs(v - outs[0]) > 1e-9:
            print("MISMATCH path", i, (c, l), "compat", compat, "err", e, v, outs[0])
print("worst relative difference staged vs whole:", worst)
