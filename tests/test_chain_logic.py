"""Chain-level arithmetic against an INDEPENDENT restatement.

The path contribution / gradient / Hessian of the oracle are pinned to the reference's compiled code (test_ref_parity.py,
test_h2mc.py).  The chain-level logic around them -- truncated drift, Adam-style moments, the `first` rule, the diagonal
Gaussian of ComputeGaussian, the asymmetric Metropolis-Hastings ratio, the augmented large-step ratio, the lastScore /
moment / counter bookkeeping -- has no reference-held vector (the reference's chain code needs Eigen + Embree and cannot be
built here), and the CPU twin the GPU is compared with is compiled from the product's own headers.  This file therefore
re-derives that logic a second time, in numpy, straight from the reference's sources
    src/mutation_mala.h:83-278   src/mala.cpp:7-51   src/gaussian.cpp:5-36   src/fastmath.h:365-381
    src/mutation_large.h:70-127  src/mlt.cpp:96-170  src/mutation_small.h:16-55
(not from csrc/core), and replays recorded chain steps through it: what each step READ (scores, moments, gradients, the
drawn offset) comes from the oracle's step recorder (lmco_chain_debug), what it PRODUCED (Gaussians, acceptance probability,
committed moments, large-step bookkeeping) must match the numpy result."""
import ctypes
import os

import numpy as np
import pytest

from conftest import SCENES

f32 = np.float32
STRIDE = 32 + 15 * 16
A = {"offset": 0, "gprop": 1, "gcur": 2, "v1b": 3, "v2b": 4, "v1a": 5, "v2a": 6, "cnv2b": 7, "pnv2b": 8,
     "cmean": 9, "cinv": 10, "ccovL": 11, "pmean": 12, "pinv": 13, "pcovL": 14}


def arr(r, name, dim):
    o = 32 + 16 * A[name]
    return r[o:o + dim].astype(f32)


def fastlog(x):
    """src/fastmath.h:365-381 (fastlog2 * ln 2), float32 arithmetic."""
    x = np.asarray(x, f32)
    i = x.view(np.uint32)
    mx = ((i & np.uint32(0x007FFFFF)) | np.uint32(0x3f000000)).view(f32)
    y = i.astype(f32) * f32(1.1920928955078125e-7)
    l2 = y - f32(124.22551499) - f32(1.498030302) * mx - f32(1.72587999) / (f32(0.3520887068) + mx)
    return f32(0.69314718) * l2


def inverse(x):
    return f32(1.0) / f32(x)


def compute_gaussian(dim, v1, M, ss, shk, sc):
    """src/mala.cpp:7-51 -> (mean, invCov_d, covL_d, logDet)"""
    shrk = inverse(f32(shk) * f32(shk))
    if sc <= f32(1e-10):
        return (np.zeros(dim, f32), np.full(dim, shrk, f32), np.full(dim, f32(shk), f32),
                f32(dim) * fastlog(inverse(f32(shk) * f32(shk))))
    cov_t = f32(ss) * f32(ss) * (M + f32(1.0))
    invcov = (f32(1.0) / cov_t + shrk).astype(f32)
    cov = (f32(1.0) / invcov).astype(f32)
    mean = (np.clip(v1, f32(-5.0), f32(5.0)) * cov / f32(2.0)).astype(f32)
    logdet = f32(0.0)
    for i in range(dim):
        logdet = f32(logdet + fastlog(invcov[i]))
    return mean, invcov, np.sqrt(cov).astype(f32), logdet


def isotropic(dim, sigma):
    """src/gaussian.cpp:5-27"""
    inv = f32(1.0) / (f32(sigma) * f32(sigma))
    return np.zeros(dim, f32), np.full(dim, inv, f32), np.full(dim, f32(sigma), f32), f32(dim) * fastlog(inv)


def mala_gaussian(opt, dim, grad, v1, v2, new_v2_prev, ssScore):
    """src/mutation_mala.h:111-129 (current) / :202-220 (proposal) -> (gaussian, new_v1, new_v2)"""
    g = grad.astype(f32).copy()
    norm = f32(0.0)
    for i in range(dim):
        norm = f32(norm + g[i] * g[i])
    norm = np.sqrt(norm).astype(f32)
    drift = f32(opt["mala-gn"])
    g = (g * (drift / max(drift, norm))).astype(f32)
    first = not (new_v2_prev[:dim] > f32(1e-10)).any()
    if first:
        nv1, nv2 = g.copy(), (g * g).astype(f32)
    else:
        nv1 = (f32(0.9) * v1 + f32(0.1) * g).astype(f32)
        nv2 = (f32(0.999) * v2 + (f32(0.001) * g) * g).astype(f32)
    M = np.clip(f32(1.0) / (f32(1e-3) + np.sqrt(nv2).astype(f32)), f32(0.01), f32(100.0)).astype(f32)
    return compute_gaussian(dim, nv1, M, opt["mala-stepsize"], opt["malastddev"], f32(ssScore)), nv1, nv2


def log_pdf(x, gauss):
    """src/gaussian.cpp:24-36 (diagonal branch)"""
    mean, invcov, _, logdet = gauss
    d = (x - mean).astype(f32)
    lp = f32(len(x)) * f32(-0.9189385332046727) + f32(0.5) * f32(logdet)
    q = f32(0.0)
    for i in range(len(x)):
        q = f32(q + d[i] * (invcov[i] * d[i]))
    return f32(lp - f32(0.5) * q)


def close(a, b, rtol=2e-5, atol=1e-30):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def run_debug(oracle, xml, opts, chain_ids, steps, chains=256):
    h = oracle.load(xml)
    for k, v in opts.items():
        oracle.set_option(h, k, v)
    norm, ls = oracle.mlt_init(h, 100000, chains, 32)
    out = []
    oracle.L.lmco_chain_debug.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                          ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    for cid in chain_ids:
        rec = np.zeros((steps, STRIDE), np.float32)
        assert oracle.L.lmco_chain_debug(h, cid, chains, steps, steps, norm, oracle.p(ls), oracle.p(rec)) == 0
        out.append(rec)
    # the recorder runs the very chains lmco_run_chains runs
    _, tr, a, _ = oracle.run_chains(h, chains, steps, norm, ls, samples_per_chain=steps)
    for cid, rec in zip(chain_ids, out):
        assert np.array_equal(rec[:, 2].view(np.uint32), a[cid].view(np.uint32))
        assert np.array_equal(rec[:, 1].astype(np.uint8), (tr[cid] >> 2) & 1)
    opt = {k: oracle.get_option(h, k) for k in ("mala-gn", "mala-stepsize", "malastddev", "perturbstddev")}
    return out, opt, norm


@pytest.mark.parametrize("scene,maxdepth", [("torus", 6), ("veachdoor", 7)])
def test_mala_and_large_step_arithmetic_vs_numpy_restatement(oracle, scene, maxdepth):
    xml = os.path.join(SCENES, scene, "lmc.xml")
    recs, opt, norm = run_debug(oracle, xml, {"maxdepth": maxdepth}, list(range(0, 256, 8)), 150)
    n_mala = n_cur = n_large = n_commit = n_first = n_iso = 0
    for rec in recs:
        for r in rec:
            kind, accepted, a = int(r[0]), int(r[1]), f32(r[2])
            if r[28] != 0:
                continue                                     # outlier reset: covered by test_chain_parity
            if kind == 0:                                    # ---- large step, src/mutation_large.h:70-116
                if a > 0:                                    # (a == 0: no contribution, the proposal state is stale)
                    if r[16]:
                        prob_prop = f32(r[8]) / f32(r[10])
                        prob_last = f32(r[11]) / f32(r[12])
                        want = np.clip((f32(r[8]) * prob_last) / (f32(r[7]) * prob_prop), f32(0), f32(1))
                    else:
                        want = f32(1.0)
                    assert close(a, want), ("large a", a, want)
                    n_large += 1
                if accepted:                                 # src/mlt.cpp:128-131
                    assert r[21] == r[8] and r[22] == r[10] and r[25] == 0 and r[27] == 0
                else:
                    assert r[21] == r[11] and r[22] == r[12] and r[27] == r[26] + 1
                continue
            if kind != 2:
                if kind == 1 and a > 0:                      # isotropic mixing step: symmetric proposal, src/mutation_small.h:41-46
                    assert close(a, np.clip(f32(r[6]) / f32(r[5]), f32(0), f32(1)))
                    n_iso += 1
                continue
            # ---- MALA small step
            dim = int(r[3])
            buffered = int(r[14])
            zero = np.zeros(16, f32)
            v1b = arr(r, "v1b", 16) if buffered else zero   # !buffered: all chain vectors are zeroed first, src/mutation_mala.h:59-81
            v2b = arr(r, "v2b", 16) if buffered else zero
            cnv2 = arr(r, "cnv2b", 16) if buffered else zero
            pnv2 = arr(r, "pnv2b", 16) if buffered else zero
            cur_rec = (arr(r, "cmean", dim), arr(r, "cinv", dim), arr(r, "ccovL", dim), f32(r[19]))
            if not r[13]:                                    # current Gaussian built in this step, src/mutation_mala.h:83-131
                mode = int(r[17])
                if mode == 2:
                    want, _, _ = mala_gaussian(opt, dim, arr(r, "gcur", dim), v1b[:dim], v2b[:dim], cnv2, r[5])
                elif mode == 0:
                    want = isotropic(dim, opt["malastddev"])
                else:
                    want = None                              # ssScore <= 1e-10: the reference reuses a stale gradient (DESIGN s4)
                if want is not None:
                    for x, y in zip(cur_rec, want):
                        assert close(x, y), ("current gaussian", x, y)
                    n_cur += 1
            if not r[15]:
                assert a == 0                                # no contribution: a = 0, src/mutation_mala.h:274-276
                continue
            pdim = int(r[4])
            assert pdim == dim                               # a perturbation keeps the path class
            modep = int(r[18])
            prop_rec = (arr(r, "pmean", dim), arr(r, "pinv", dim), arr(r, "pcovL", dim), f32(r[20]))
            nv1 = nv2 = None
            if modep == 2:
                want, nv1, nv2 = mala_gaussian(opt, dim, arr(r, "gprop", dim), v1b[:dim], v2b[:dim], pnv2, r[6])
                n_first += int(not (pnv2[:dim] > 1e-10).any())
            elif modep == 0:
                want = isotropic(dim, opt["malastddev"])
            else:
                want = None
            if want is not None:
                for x, y in zip(prop_rec, want):
                    assert close(x, y), ("proposal gaussian", x, y)
            # asymmetric MH ratio, src/mutation_mala.h:262-267
            off = arr(r, "offset", dim)
            py = log_pdf(off, cur_rec)
            px = log_pdf(-off, prop_rec)
            want_a = np.clip(np.exp(np.float64(px) - np.float64(py)) * np.float64(r[6]) / np.float64(r[5]), 0.0, 1.0)
            assert abs(float(a) - want_a) <= 2e-4 * max(want_a, 1e-3), ("mala a", a, want_a)    # exp of a difference of ~1e2-sized logs
            n_mala += 1
            if accepted:                                     # moment commit, src/mlt.cpp:133-141
                assert r[24] == r[23] + 1 and r[25] == 1 and r[27] == 0
                if nv1 is not None:
                    assert close(arr(r, "v1a", dim), nv1) and close(arr(r, "v2a", dim), nv2)
                    n_commit += 1
            else:
                assert r[24] == r[23] and r[27] == r[26] + 1
                if buffered:
                    assert np.array_equal(arr(r, "v1a", dim), arr(r, "v1b", dim))
    assert n_mala > 1500 and n_cur > 50 and n_large > 100 and n_commit > 500 and n_first > 30 and n_iso > 100, \
        (n_mala, n_cur, n_large, n_commit, n_first, n_iso)
