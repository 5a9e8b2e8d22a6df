"""The staged (per-vertex wavefront) forms of PerturbPathBidir / GeneratePathBidir
(langevin-mcmc_b200/csrc/core/stages.h) against the monolithic restatement (core/path.h, following
src/path.cpp:1237-1449 and :1953-2160): same chains, bit for bit.  CPU only -- the device runs the
very same stage functions from its shading kernels, with the ray queries in between done by the
traversal kernels (tests/test_chain_parity.py compares that with the host twin on the GPU)."""
import numpy as np
import pytest


def _run(oracle, xml, opts, chains, steps, staged, n_init=60000):
    h = oracle.load(xml)
    for k, v in opts.items():
        oracle.set_option(h, k, v)
    norm, ls = oracle.mlt_init(h, n_init, chains, 32)
    oracle.use_staged(staged)
    try:
        return oracle.run_chains(h, chains, steps, norm, ls, threads=8)
    finally:
        oracle.use_staged(False)


@pytest.mark.parametrize("scene,opts,chains,steps", [
    ("torus", {"maxdepth": 4}, 512, 100),
    ("torus", {"maxdepth": 8}, 256, 64),
    ("door", {"maxdepth": 12}, 128, 48),
    ("torus", {"maxdepth": 8, "h2mc": 1, "mala": 0}, 96, 32),
    ("torus", {"maxdepth": 6, "mala": 0}, 128, 64),
])
def test_staged_equals_monolithic(oracle, torus_xml, door_xml, scene, opts, chains, steps):
    xml = torus_xml if scene == "torus" else door_xml
    import ctypes
    oracle.L.lmco_deferred_overflows.restype = ctypes.c_long
    dropped0 = oracle.L.lmco_deferred_overflows()
    film0, tr0, a0, st0 = _run(oracle, xml, opts, chains, steps, False)
    film1, tr1, a1, st1 = _run(oracle, xml, opts, chains, steps, True)
    # capacity invariant of DeferredList (core/stages.h): no candidate was ever dropped for lack of a slot
    assert oracle.L.lmco_deferred_overflows() == dropped0
    assert np.array_equal(tr0, tr1)
    assert np.array_equal(a0.view(np.uint32), a1.view(np.uint32))
    assert np.array_equal(st0, st1)
    # per-thread films are summed in scheduling order: equal up to fp32 summation order
    assert np.allclose(film0, film1, rtol=1e-4, atol=1e-6)
    assert (tr0 & 3 == 0).any() and (tr0 & 3 != 0).any()    # large and small steps both exercised


def test_deferred_list_equals_immediate_vector_on_scripted_events(oracle):
    """DeferredList (visibility answered later, contributions kept in push order, conditional clears) against
    the reference's immediate std::vector semantics, on random event scripts including the clear() cases that
    real paths almost never produce (src/path.cpp:700-704, 1345-1347)."""
    import ctypes
    rng = np.random.default_rng(7)
    for trial in range(400):
        n = int(rng.integers(1, 40))
        types = rng.choice([0, 1, 1, 1, 2, 3], size=n, p=[0.2, 0.25, 0.25, 0.2, 0.07, 0.03]).astype(np.int32)
        occl = rng.integers(0, 2, size=n).astype(np.int32)
        imm = np.zeros(64, np.float32); dfr = np.zeros(64, np.float32)
        ni = ctypes.c_int(); nd = ctypes.c_int()
        rc = oracle.L.lmco_deferred_probe(n, oracle.p(types), oracle.p(occl), oracle.p(imm), ctypes.byref(ni), oracle.p(dfr), ctypes.byref(nd))
        assert rc == 0
        assert ni.value == nd.value and np.array_equal(imm[:ni.value], dfr[:nd.value]), (types, occl)
