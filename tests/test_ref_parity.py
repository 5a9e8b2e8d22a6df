"""Oracle (and, on the GPU, the CUDA evaluator) vs the REFERENCE'S OWN generated path functions.

Golden vectors: tests/golden/path_golden.npz (made by tests/golden/make_path_golden.py from
oracle/_ref, i.e. the reference's src/bin/*.c / *.ispc compiled unmodified).  When oracle/_ref is
present the live libraries are exercised too.

Tolerances (north star: contribution / gradient within 1e-4 relative):
  * forward value log(Luminance(contrib)): |ours - ref| <= 1e-4 absolute on the log == 1e-4
    relative on the contribution, for >= 99% of the paths; the remainder are the reference's
    6-decimal constants amplified by near-singular configurations and must
    stay below 5e-3.
  * gradient, option adjointcompat = 1 (default: one reverse sweep in the reference's merge order,
    csrc/core/pathgrad_rev.h) vs the reference's reverse-mode code (MALA library) on EVERY golden
    path -- including paths through RoughDielectric vertices and the light-tracing classes
    (camDepth == 1), where the reference's sweep deviates from the true gradient by 1-100 %
    (chad merges `if` outputs by assignment): relative L2 error <= 1e-4 median,
    <= 5e-3 at the 99th percentile in each bucket.  (Both sides are fp32; the reference is built with
    ispc fast-math and prints its constants with 6 decimals: the float64 interpretation of the same
    program, tools/adgen/check_whole.py, sits at median 2e-6 / p99 1.3e-3 against it.)
  * gradient, adjointcompat = 0 (reverse sweep, true adjoint) and = 2 (forward-mode duals,
    csrc/core/pathgrad.h) vs the reference's forward-mode code (H2MC library, every path length up to 8):
    <= 1e-4 median, <= 5e-3 p99; and against each other <= 1e-5 median (two independent derivations of the same
    gradient).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bsdf_types


def point_xml():
    """torus with the point emitter the reference keeps commented out in its scene file (scene id 2 of the fixture)."""
    return os.path.join(ROOT, "scenes", "torus", "point.xml")


def load_golden():
    g = np.load(os.path.join(GOLDEN, "path_golden.npz"))
    return {k: g[k] for k in g.files}


def rel_l2(a, b):
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-3)


def evaluate_golden(eval_fn, g):
    """eval_fn(scene_id, c, l, primary[n, D+1], vert[n, VS]) -> (loglum[n], grad[n, D])"""
    n = len(g["c"])
    ll = np.zeros(n, np.float32)
    grads = [None] * n
    keys = sorted(set(zip(g["scene"].tolist(), g["c"].tolist(), g["l"].tolist())))
    for (s, c, l) in keys:
        idx = np.where((g["scene"] == s) & (g["c"] == c) & (g["l"] == l))[0]
        dim = 2 * max(c + l - 1, 2)
        prim = np.ascontiguousarray(g["primary"][idx, :dim + 1])
        vert = np.ascontiguousarray(g["vert"][idx])
        a, b = eval_fn(s, c, l, prim, vert)
        ll[idx] = a
        for k, i in enumerate(idx):
            grads[i] = b[k]
    return ll, grads


def check_against_golden(ll, grads, g, mode=1):
    n = len(ll)
    # The mutation only evaluates the function when ssScore > 1e-10 (src/mutation_mala.h:100);
    # below that fp32 products underflow while the reference's C forward code runs in double.
    gate = g["ss"] > 1e-10
    fwd_ok = np.isfinite(g["ref_fwd"]) & gate
    assert np.array_equal(np.isfinite(ll)[gate], fwd_ok[gate]), "finite / non-finite pattern of the forward value differs"
    d = np.abs(ll[fwd_ok] - g["ref_fwd"][fwd_ok])
    assert np.percentile(d, 99) <= 1e-4 * 5, "forward value p99 %g" % np.percentile(d, 99)
    assert (d <= 1e-4).mean() >= 0.97, "forward value within 1e-4: %.3f" % (d <= 1e-4).mean()
    assert d.max() <= 5e-3, "forward value max %g" % d.max()
    e_fm, e_rev_noglass, e_rev_glass = [], [], []
    for i in range(n):
        c, l = int(g["c"][i]), int(g["l"][i])
        dim = 2 * max(c + l - 1, 2)
        if not fwd_ok[i] or not np.isfinite(grads[i]).all():
            continue
        fm, rv = g["ref_fwdm"][i, :dim], g["ref_rev"][i, :dim]
        if np.isfinite(fm).all():
            e_fm.append(rel_l2(grads[i], fm))
        if np.isfinite(rv).all():
            buggy = (2 in bsdf_types(c, l, g["vert"][i])) or c == 1
            (e_rev_glass if buggy else e_rev_noglass).append(rel_l2(grads[i], rv))
    e_fm, e_rev_noglass, e_rev_glass = map(np.array, (e_fm, e_rev_noglass, e_rev_glass))
    assert len(e_fm) > 100 and len(e_rev_noglass) > 100 and len(e_rev_glass) > 100
    rep = dict(fwd_med=float(np.median(d)), fwd_max=float(d.max()), fm_med=float(np.median(e_fm)),
               fm_p99=float(np.percentile(e_fm, 99)), rev_noglass_med=float(np.median(e_rev_noglass)),
               rev_noglass_p99=float(np.percentile(e_rev_noglass, 99)),
               rev_glass_med=float(np.median(e_rev_glass)), rev_glass_p90=float(np.percentile(e_rev_glass, 90)),
               rev_glass_p99=float(np.percentile(e_rev_glass, 99)), n_glass=len(e_rev_glass))
    if mode == 1:     # the reference's reverse sweep, every bucket
        assert rep["rev_noglass_med"] <= 1e-4 and rep["rev_noglass_p99"] <= 5e-3, rep
        assert rep["rev_glass_med"] <= 1e-4 and rep["rev_glass_p99"] <= 5e-3, rep
    else:             # the true gradient
        assert rep["fm_med"] <= 1e-4 and rep["fm_p99"] <= 5e-3, rep
        assert rep["rev_noglass_med"] <= 1e-4 and rep["rev_noglass_p99"] <= 5e-3, rep
    return rep


def test_oracle_evaluator_matches_reference_golden(oracle, torus_xml, door_xml):
    g = load_golden()
    handles = {0: oracle.load(torus_xml), 1: oracle.load(door_xml), 2: oracle.load(point_xml())}
    # the scene block the golden inputs were evaluated with must be what the loader produces
    for s, h in handles.items():
        idx = np.where(g["scene"] == s)[0][0]
        assert np.allclose(oracle.scene_serialized(h), g["scene_ser"][idx], rtol=2e-6, atol=1e-6)
    by_mode = {}
    for mode in (1, 0, 2):
        for h in handles.values():
            oracle.set_option(h, "adjointcompat", mode)
        ll, grads = evaluate_golden(lambda s, c, l, p, v: oracle.eval_batch(handles[s], c, l, p, v), g)
        rep = check_against_golden(ll, grads, g, mode)
        by_mode[mode] = grads
        print("oracle (adjointcompat=%d) vs reference golden:" % mode, rep)
    # reverse sweep (true adjoint) == forward-mode duals: two independent derivations
    e = [rel_l2(a, b) for a, b in zip(by_mode[0], by_mode[2]) if np.isfinite(a).all() and np.isfinite(b).all()]
    assert np.median(e) <= 1e-5 and np.percentile(e, 99) <= 5e-3, (np.median(e), np.percentile(e, 99))
    # the tracer's own score must agree with the AD twin wherever the reference's twin is sane
    ok = np.isfinite(ll) & (g["ss"] > 1e-30)
    d = np.abs(ll[ok] - np.log(g["ss"][ok]))
    assert np.median(d) < 1e-4   # the AD twin adds epsilon guards the tracer does not have (App. B#5)


def test_live_reference_library_agrees_with_golden(oracle, ref_mala):
    """oracle/_ref built here must reproduce the committed vectors (guards the fixture)."""
    g = load_golden()
    for i in range(0, len(g["c"]), 7):
        c, l = int(g["c"][i]), int(g["l"][i])
        out = np.zeros(1, np.float32)
        f = getattr(ref_mala, "evaluate_path_bidir_mala_%d_%d_static" % (c, l))
        vert = np.zeros(1005, np.float32)
        vert[:g["vert"].shape[1]] = g["vert"][i]
        prim = np.zeros(25, np.float32)
        prim[:17] = g["primary"][i]
        f(oracle.p(np.ascontiguousarray(g["lens"][i])), oracle.p(prim), oracle.p(np.ascontiguousarray(g["scene_ser"][i])),
          oracle.p(vert), oracle.p(out))
        assert (np.isnan(out[0]) and np.isnan(g["ref_fwd"][i])) or out[0] == g["ref_fwd"][i]


@pytest.mark.gpu
def test_cuda_eval_batch_matches_reference_golden_and_oracle(lmc, oracle, torus_xml, door_xml):
    g = load_golden()
    scenes = {0: lmc.ParseScene(torus_xml), 1: lmc.ParseScene(door_xml), 2: lmc.ParseScene(point_xml())}
    ctxs = {s: lmc.ChainContext(sc, 0) for s, sc in scenes.items()}
    handles = {0: oracle.load(torus_xml), 1: oracle.load(door_xml), 2: oracle.load(point_xml())}

    def gpu_eval(s, c, l, p, v):
        lens = np.zeros((len(p), 2), np.float32)
        return ctxs[s].eval_batch(c, l, lens, p, v)

    ll, grads = evaluate_golden(gpu_eval, g)
    rep = check_against_golden(ll, grads, g)
    print("cuda vs reference golden:", rep)
    # and bit-for-bit against the CPU twin (same statements, fmad off / contraction off)
    ll_o, grads_o = evaluate_golden(lambda s, c, l, p, v: oracle.eval_batch(handles[s], c, l, p, v), g)
    def bit_equal(a, b):   # NaN payloads differ between x86 and sm_100; NaN == NaN here
        a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
        return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))

    assert bit_equal(ll, ll_o)
    nbad = sum(0 if bit_equal(a, b) else 1 for a, b in zip(grads, grads_o))
    assert nbad == 0, "%d of %d gradients differ bitwise between CUDA and the CPU twin" % (nbad, len(grads))


def test_point_light_paths_match_reference_code(oracle):
    """PointLight (src/pointlight.cpp:37-116) is in neither bundled scene's active configuration; scene 2 of the
    fixture enables the point emitter of scenes/torus/lmc.xml.  Forward value and reverse-mode gradient of its
    direct-lighting (c, 1) and light-subpath (c, >= 2) classes against the reference's compiled code, incl. the
    twin's `lightType == PointLight` MIS term."""
    g = load_golden()
    h = oracle.load(point_xml())
    idx = np.where(g["scene"] == 2)[0]
    assert len(idx) > 50 and (g["l"][idx] >= 2).any() and (g["l"][idx] == 1).any()
    errs, fwd = [], []
    for i in idx:
        c, l = int(g["c"][i]), int(g["l"][i])
        dim = 2 * max(c + l - 1, 2)
        # light type of the record really is PointLight (0): direct-light record or emitting light
        ll, gr = oracle.eval_batch(h, c, l, np.ascontiguousarray(g["primary"][i:i + 1, :dim + 1]), np.ascontiguousarray(g["vert"][i:i + 1]))
        if not (np.isfinite(g["ref_fwd"][i]) and g["ss"][i] > 1e-10 and np.isfinite(g["ref_rev"][i, :dim]).all()):
            continue
        fwd.append(abs(ll[0] - g["ref_fwd"][i]))
        errs.append(rel_l2(gr[0], g["ref_rev"][i, :dim]))
    errs, fwd = np.array(errs), np.array(fwd)
    assert len(errs) > 40
    assert np.median(fwd) <= 1e-5 and fwd.max() <= 5e-3, (np.median(fwd), fwd.max())
    assert np.median(errs) <= 1e-4 and np.percentile(errs, 95) <= 5e-3, (np.median(errs), np.percentile(errs, 95))
