"""H2MC mutation (BASELINE configs[3]): Hessian by forward-over-reverse (second-order forward mode
as the cross-check), parallel-order Jacobi eigen-solver (replacing Eigen::SelfAdjointEigenSolver, unpinned; warp-cooperative
on the device), dense Gaussian proposal."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES, bsdf_types


def test_jacobi_eigensolver_matches_lapack(oracle):
    rng = np.random.default_rng(3)
    for n in (2, 3, 6, 8, 12, 16):
        for scale in (1.0, 1e4):
            M = rng.normal(size=(n, n)).astype(np.float32)
            A = np.ascontiguousarray((M + M.T) * np.float32(scale))
            V = np.zeros((n, n), np.float32)
            w = np.zeros(n, np.float32)
            assert oracle.L.lmco_jacobi(n, oracle.p(A), oracle.p(V), oracle.p(w)) == 0
            we = np.linalg.eigvalsh(A.astype(np.float64))
            assert np.all(np.diff(w) >= 0)                                   # ascending
            assert np.abs(w - we).max() <= 2e-5 * np.abs(we).max()
            assert np.abs(V.T @ V - np.eye(n)).max() < 2e-5                  # orthonormal
            assert np.abs((V * w) @ V.T - A).max() <= 3e-5 * np.abs(A).max()  # A = V diag(w) V^T
            lead = np.array([V[np.nonzero(V[:, j])[0][0], j] for j in range(n)])
            assert (lead > 0).all()                                          # sign convention


@pytest.mark.parametrize("mode", [1, 3])
def test_hessian_matches_reference_generated_code(oracle, torus_xml, door_xml, mode):
    """Our Hessians vs the reference's (evaluate_path_bidir_<c>_<l>_static_derv, vectors in
    tests/golden/path_golden.npz, every path length up to 8):
      mode 3  forward-over-reverse: the generated reverse sweep on dual numbers (pathgrad_rev.h path_loglum_hess_rev)
              -- what the H2MC mutation runs, on the device and in the host twin
      mode 1  second-order forward mode (csrc/core/pathgrad.h path_loglum_hess) -- an independent derivation kept as
              a cross-check (it was the device path of round 1)."""
    g = np.load(os.path.join(GOLDEN, "path_golden.npz"))
    handles = {0: oracle.load(torus_xml), 1: oracle.load(door_xml), 2: oracle.load(os.path.join(SCENES, "torus", "point.xml"))}
    for h in handles.values():
        oracle.set_option(h, "adjointcompat", mode)
    errs, errs_glass, lens = [], [], []
    for i in range(len(g["c"])):
        c, l, s = int(g["c"][i]), int(g["l"][i]), int(g["scene"][i])
        dim = 2 * max(c + l - 1, 2)
        if c + l - 1 > 8 or g["ss"][i] <= 1e-10 or not np.isfinite(g["ref_hess"][i, :dim * dim]).all():
            continue
        prim = np.ascontiguousarray(g["primary"][i:i + 1, :dim + 1])
        vert = np.ascontiguousarray(g["vert"][i:i + 1])
        ll = np.zeros(1, np.float32)
        og = np.zeros(dim, np.float32)
        oh = np.zeros(dim * dim, np.float32)
        assert oracle.L.lmco_eval_batch_hess(handles[s], c, l, 1, oracle.p(prim), dim + 1, oracle.p(vert), vert.shape[1],
                                             oracle.p(ll), oracle.p(og), oracle.p(oh)) == 0
        if not np.isfinite(oh).all():
            continue
        H = g["ref_hess"][i, :dim * dim].reshape(dim, dim)
        O = oh.reshape(dim, dim)
        assert np.array_equal(O, O.T)                  # ours is symmetric by construction (mode 3: symmetrised)
        e = np.abs(O - H).max() / (np.abs(H).max() + 1e-3)
        (errs_glass if 2 in bsdf_types(c, l, g["vert"][i]) else errs).append(e)
        lens.append(c + l - 1)
        assert np.abs(og - g["ref_fwdm"][i, :dim]).max() <= 5e-2 * (np.abs(og).max() + 1e-3)   # sanity only; bounds in test_ref_parity
    errs, errs_glass = np.array(errs), np.array(errs_glass)
    assert len(errs) > 150 and max(lens) == 8 and sum(1 for x in lens if x >= 6) > 60   # path lengths 6-8 are pinned too (H2MC at maxdepth 8)
    assert np.median(errs) <= 1e-4 and np.percentile(errs, 90) <= 2e-3, (np.median(errs), np.percentile(errs, 90))
    # glass: the reference's reverse half carries its merge bug; report, bound loosely
    print("hessian rel err: no glass med %.2e p90 %.2e | glass med %.2e p90 %.2e" %
          (np.median(errs), np.percentile(errs, 90), np.median(errs_glass), np.percentile(errs_glass, 90)))
    assert np.median(errs_glass) <= 2e-2


def test_h2mc_oracle_chain_properties(oracle):
    h = oracle.load(os.path.join(SCENES, "torus", "h2mc.xml"))
    oracle.set_option(h, "maxdepth", 6)
    norm, ls = oracle.mlt_init(h, 60000, 256, 32)
    f, t, a, s = oracle.run_chains(h, 256, 40, norm, ls, samples_per_chain=40, threads=4)
    types = t & 3
    assert (types != 3).all()                 # no MALA steps in an h2mc run
    assert (types == 2).mean() > 0.4          # H2MC small steps dominate (largestepprob 0.2, 10% isotropic)
    assert s[8] > 0 and np.isfinite(f).all() and f.sum() > 0
    assert 0.2 < s[6] / s[2] < 0.9            # acceptance rate of the anisotropic proposals


@pytest.mark.gpu
def test_cuda_h2mc_chain_bit_identical_to_oracle(lmc, oracle):
    xml = os.path.join(SCENES, "torus", "h2mc.xml")
    for maxdepth, chains, steps in ((5, 512, 40), (8, 256, 24)):
        sc = lmc.ParseScene(xml)
        assert sc.options["h2mc"] == 1
        sc.options["maxdepth"] = maxdepth
        norm, init_ls = lmc.MLTInit(sc, 200000, chains, 32)
        ctx = lmc.ChainContext(sc, 0)
        ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
        trace, a = ctx.run(steps, trace=True, a_trace=True)
        st = ctx.stats()
        h = oracle.load(xml)
        oracle.set_option(h, "maxdepth", maxdepth)
        ofilm, otrace, oa, ostats = oracle.run_chains(h, chains, steps, norm, init_ls, samples_per_chain=steps)
        assert np.array_equal(trace, otrace), "%d chains diverge" % int((trace != otrace).any(axis=1).sum())
        assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))
        assert st["gradient_evals"] == int(ostats[8])
        assert np.allclose(ctx.film(), ofilm, rtol=1e-4, atol=1e-5 * max(1.0, float(ofilm.max())))


@pytest.mark.gpu
def test_cuda_hessian_bit_equal_to_oracle(lmc, oracle, torus_xml):
    g = np.load(os.path.join(GOLDEN, "path_golden.npz"))
    ctx = lmc.ChainContext(lmc.ParseScene(torus_xml), 0)
    h = oracle.load(torus_xml)
    for (c, l) in ((3, 1), (4, 0), (5, 1)):
        idx = np.where((g["scene"] == 0) & (g["c"] == c) & (g["l"] == l))[0]
        dim = 2 * max(c + l - 1, 2)
        prim = np.ascontiguousarray(g["primary"][idx, :dim + 1])
        vert = np.ascontiguousarray(g["vert"][idx])
        ll, gr, he = ctx.eval_batch(c, l, np.zeros((len(idx), 2), np.float32), prim, vert, want_hess=True)
        ol = np.zeros(len(idx), np.float32)
        og = np.zeros((len(idx), dim), np.float32)
        oh = np.zeros((len(idx), dim, dim), np.float32)
        oracle.L.lmco_eval_batch_hess(h, c, l, len(idx), oracle.p(prim), dim + 1, oracle.p(vert), vert.shape[1], oracle.p(ol),
                                      oracle.p(og), oracle.p(oh))
        for a_, b_ in ((ll, ol), (gr, og), (he, oh)):
            assert np.array_equal(np.isnan(a_), np.isnan(b_))
            assert np.array_equal(a_[~np.isnan(a_)].view(np.uint32), b_[~np.isnan(b_)].view(np.uint32))
