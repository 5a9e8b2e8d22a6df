#!/usr/bin/env python3
"""Generate tests/golden/path_golden.npz: serialized paths (the reference's PathFunc ABI inputs,
src/path.h:121-125) recorded from both bundled scenes, together with the outputs of the
REFERENCE's OWN generated code compiled into oracle/_ref/ (oracle/build_ref.sh):

  ref_fwd   evaluate_path_bidir_mala_<c>_<l>_static        log Luminance(contrib)
  ref_rev   evaluate_path_bidir_mala_<c>_<l>_static_derv   reverse-mode gradient (what LMC uses)
  ref_fwdm  evaluate_path_bidir_<c>_<l>_static_derv        forward-mode gradient (H2MC library)
  ref_hess  the same call's Hessian, hess[i * D + j] = d(grad_j)/dx_i, every class (c + l - 1 <= 8, D <= 16)

Scenes: 0 = torus (environment light), 1 = veach-door (area light), 2 = torus with the point emitter the reference
keeps commented out in its scene file (scenes/torus/point.xml; the only PointLight configuration).

The inputs are produced by the oracle's path recorder (lmco_sample_paths: GeneratePathBidir ->
ToSubpath -> optional PerturbPathBidir -> Serialize).  For (c, 0) paths that end on the
environment map the reference leaves the last shape block of its reused buffer stale
(src/path.cpp:2547-2550); an all-zero block makes its reverse sweep emit NaN, so the recorder's
zero block is replaced by a copy of the first vertex's triangle ("miss_patch"), which leaves the
forward value unchanged.  Run in the build container:  python tests/golden/make_path_golden.py
"""
import collections
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import Oracle, ROOT, ref_lib  # noqa: E402

VS = 700   # trimmed vertParams stride kept in the fixture (classes up to c + l - 1 = 8 need <= 605)
PER_CLASS = 12


def main():
    o = Oracle()
    mala, hess = ref_lib("libpathref_mala.so"), ref_lib("libpathref_hess.so")
    assert mala is not None and hess is not None, "run `make ref` first"
    recs = []
    for sid, (scene, xml) in enumerate((("torus", "lmc.xml"), ("veachdoor", "lmc.xml"), ("torus", "point.xml"))):
        h = o.load(os.path.join(ROOT, "scenes", scene, xml))
        ser = o.scene_serialized(h)
        out = o.sample_paths(h, seed=2024, num_large_steps=3000, perturb=True, max_len=8, max_records=40000)
        cl = out[:, :2].astype(int)
        per = collections.Counter()
        for i in range(len(out)):
            c, l = int(cl[i, 0]), int(cl[i, 1])
            if per[(c, l)] >= PER_CLASS:
                continue
            per[(c, l)] += 1
            prim = np.ascontiguousarray(out[i, 16:41])
            vert = np.ascontiguousarray(out[i, 41:41 + Oracle.VSTRIDE]).copy()
            if l == 0:
                off = 3 + (c - 2) * 59
                if not vert[off:off + 46].any():
                    vert[off:off + 46] = vert[3:3 + 46]     # miss_patch
            dim = 2 * max(c + l - 1, 2)
            lens = np.ascontiguousarray(out[i, 2:4])
            fwd = np.zeros(1, np.float32)
            rev = np.zeros(dim, np.float32)
            getattr(mala, "evaluate_path_bidir_mala_%d_%d_static" % (c, l))(o.p(lens), o.p(prim), o.p(ser), o.p(vert), o.p(fwd))
            getattr(mala, "evaluate_path_bidir_mala_%d_%d_static_derv" % (c, l))(o.p(lens), o.p(prim), o.p(ser), o.p(vert), o.p(rev), None)
            fwdm = np.full(16, np.nan, np.float32)
            rhess = np.full(256, np.nan, np.float32)     # reference Hessian (row = direction), dim <= 16
            fh = getattr(hess, "evaluate_path_bidir_%d_%d_static_derv" % (c, l), None) if c + l - 1 <= 8 else None
            if fh is not None:
                g = np.zeros(dim, np.float32)
                hh = np.zeros(dim * dim, np.float32)
                fh(o.p(lens), o.p(prim), o.p(ser), o.p(vert), o.p(g), o.p(hh))
                fwdm[:dim] = g
                rhess[:dim * dim] = hh
            rv = np.full(16, np.nan, np.float32)
            rv[:dim] = rev
            recs.append(dict(scene=sid, c=c, l=l, lens=lens, primary=prim[:17], vert=vert[:VS],
                             scene_ser=ser, ls=out[i, 4], ss=out[i, 5], ref_fwd=fwd[0], ref_rev=rv, ref_fwdm=fwdm, ref_hess=rhess))
        print(scene, xml, "classes", sorted(per.items()))
    keys = recs[0].keys()
    arrays = {k: np.stack([np.asarray(r[k]) for r in recs]) for k in keys}
    np.savez_compressed(os.path.join(HERE, "path_golden.npz"), **arrays)
    print("wrote", len(recs), "records")


if __name__ == "__main__":
    main()
