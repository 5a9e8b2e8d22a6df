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
    film0, tr0, a0, st0 = _run(oracle, xml, opts, chains, steps, False)
    film1, tr1, a1, st1 = _run(oracle, xml, opts, chains, steps, True)
    assert np.array_equal(tr0, tr1)
    assert np.array_equal(a0.view(np.uint32), a1.view(np.uint32))
    assert np.array_equal(st0, st1)
    # per-thread films are summed in scheduling order: equal up to fp32 summation order
    assert np.allclose(film0, film1, rtol=1e-4, atol=1e-6)
    assert (tr0 & 3 == 0).any() and (tr0 & 3 != 0).any()    # large and small steps both exercised
