"""Whole-path programs recorded by pathfn.py, interpreted numerically (float64), against the reference's own
compiled code for the same serialized inputs (tests/golden/path_golden.npz: ref_fwd = generated C forward
function, ref_rev = generated ISPC reverse-mode gradient, ref_fwdm = forward-mode gradient of the Hessian library)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import chadlike as cl
import pathfn as pf

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(c, l):
    cl.begin_function()
    params = pf.Params()
    dim = 2 * max(c + l - 1, 2)
    pss = [cl.inp("primary[%d]" % (i + 1), True) for i in range(dim)]
    out = pf.record_path_function(c, l, params, pss)
    f = cl.end_function()
    inputs = dict(params.nodes)
    for n in pss:
        inputs[n.name] = n
    return cl.Program(f, inputs, [out]), dim


def rel_l2(a, b):
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-3)


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "path_golden.npz"))
    progs = {}
    rows = []
    sel = range(len(g["c"])) if len(sys.argv) < 2 else range(0, len(g["c"]), int(sys.argv[1]))
    for i in sel:
        c, l = int(g["c"][i]), int(g["l"][i])
        if (c, l) not in progs:
            progs[(c, l)] = build(c, l)
        prog, dim = progs[(c, l)]
        vals = {}
        for name in prog.inputs:
            arr, idx = name[:-1].split("[")
            idx = int(idx)
            src = {"primary": g["primary"][i], "scene": g["scene_ser"][i], "vert": g["vert"][i]}[arr]
            vals[name] = float(src[idx])
        outs, adj = prog.evaluate(vals, [1.0], compat=True)
        _, adj_exact = prog.evaluate(vals, [1.0], compat=False)
        grad = np.array([adj["primary[%d]" % (k + 1)] for k in range(dim)])
        grad_x = np.array([adj_exact["primary[%d]" % (k + 1)] for k in range(dim)])
        rows.append((i, c, l, outs[0], float(g["ref_fwd"][i]), rel_l2(grad, g["ref_rev"][i, :dim]),
                     rel_l2(grad_x, g["ref_fwdm"][i, :dim]), rel_l2(grad_x, g["ref_rev"][i, :dim]), float(g["ss"][i])))
    rows = np.array(rows)
    ok = np.isfinite(rows[:, 4]) & (rows[:, 8] > 1e-10) & np.isfinite(rows[:, 5])
    print("paths", len(rows), "usable", ok.sum())
    print("forward |dlog| median %.2e max %.2e" % (np.median(np.abs(rows[ok, 3] - rows[ok, 4])), np.max(np.abs(rows[ok, 3] - rows[ok, 4]))))
    for name, col in (("compat vs ref reverse", 5), ("exact vs ref forward-mode", 6), ("exact vs ref reverse", 7)):
        e = rows[ok, col]
        e = e[np.isfinite(e)]
        print("%-28s median %.2e p90 %.2e p99 %.2e max %.2e  (n=%d)" % (name, np.median(e), np.percentile(e, 90), np.percentile(e, 99), e.max(), len(e)))
    bad = rows[ok][rows[ok, 5] > 1e-3]
    print("compat mismatches > 1e-3:", len(bad))
    for r in bad[:30]:
        print("  path %d class (%d,%d) compat-err %.3e exact-vs-fwdm %.3e exact-vs-rev %.3e" % (r[0], r[1], r[2], r[5], r[6], r[7]))


if __name__ == "__main__":
    main()
