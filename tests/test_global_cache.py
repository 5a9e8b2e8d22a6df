"""Global cache: the cross-chain cache of adaptation states, src/global_cache.h:16-163, filled at
accepted large steps (src/mlt.cpp:121-127) and consulted instead of the gradient once a dimension holds
PSS_MAX_SIZE = 3000 entries (src/mutation_mala.h:131-161).  Option `globalcache` (default 0 = the benchmark mode).
The reference's fill order depends on thread timing; here it is defined (chain order, iteration by iteration), so the
CPU oracle and the GPU agree bit for bit."""
import os

import numpy as np
import pytest

from conftest import SCENES


def numpy_query(entries, dim, pss):
    """Restatement of global_cache_t::query with our documented tie rule (first 5 in-radius entries in insertion order)."""
    e = entries.reshape(3000, 3, dim)
    d = ((e[:, 0, :] - pss[None, :]) ** 2).astype(np.float32).sum(axis=1, dtype=np.float32)
    idx = np.where(d < np.float32(dim) * np.float32(0.01) ** 2)[0][:5]
    if len(idx) == 0:
        return None
    w = 1.0 / (d[idx].astype(np.float64) ** 2 + 1e-6)
    v1 = (e[idx, 1, :] * w[:, None]).sum(axis=0) / w.sum()
    v2 = (e[idx, 2, :] * w[:, None]).sum(axis=0) / w.sum()
    return v1, v2


def test_grid_query_equals_linear_scan(oracle):
    """The query grid over the first three PSS coordinates (core/scene.h; the reference walks a KD-tree) returns exactly
    what the linear scan returns: clustered entries so that many queries have 1 .. > 5 neighbours, queries on cell borders
    and outside [0, 1)."""
    rng = np.random.default_rng(5)
    for dim in (6, 10, 12):
        centres = rng.uniform(0.0, 1.0, size=(40, dim)).astype(np.float32)
        entries = rng.uniform(0, 1, size=(3000, 3, dim)).astype(np.float32)
        for i in range(1500):                                     # half of the entries in tight clusters
            entries[i, 0, :] = centres[i % 40] + rng.normal(size=dim).astype(np.float32) * np.float32(0.006)
        flat = np.ascontiguousarray(entries.reshape(-1))
        hits = 0
        queries = [centres[k % 40] + rng.normal(size=dim).astype(np.float32) * np.float32(0.005) for k in range(200)]
        queries += [np.round(centres[k] * 24).astype(np.float32) / np.float32(24) for k in range(20)]      # on cell borders
        queries += [np.full(dim, -0.01, np.float32), np.full(dim, 1.01, np.float32)]
        for q in queries:
            q = np.ascontiguousarray(q, np.float32)
            out = {}
            for grid in (1, 0):
                oracle.L.lmco_cache_use_grid(grid)
                v1 = np.zeros(dim, np.float32); v2 = np.zeros(dim, np.float32)
                rc = oracle.L.lmco_cache_query(dim, oracle.p(flat), oracle.p(q), oracle.p(v1), oracle.p(v2))
                out[grid] = (rc, v1, v2)
            oracle.L.lmco_cache_use_grid(1)
            assert out[0][0] == out[1][0]
            assert np.array_equal(out[0][1].view(np.uint32), out[1][1].view(np.uint32))
            assert np.array_equal(out[0][2].view(np.uint32), out[1][2].view(np.uint32))
            hits += out[1][0]
        assert hits > 100


@pytest.mark.parametrize("dim", [6, 8, 12])
def test_cache_query_semantics(oracle, dim):
    rng = np.random.default_rng(dim)
    entries = rng.uniform(0, 1, size=(3000, 3, dim)).astype(np.float32)
    centre = rng.uniform(0.2, 0.8, dim).astype(np.float32)
    # 9 entries inside the query radius (sqrt(dim) * 0.01), scattered through the buffer; the rest far away
    near = [2500, 17, 1200, 400, 2999, 90, 1800, 5, 700]
    for k, i in enumerate(near):
        entries[i, 0, :] = centre + rng.uniform(-1, 1, dim).astype(np.float32) * np.float32(0.004)
    v1 = np.zeros(dim, np.float32)
    v2 = np.zeros(dim, np.float32)
    flat = np.ascontiguousarray(entries.reshape(-1))
    assert oracle.L.lmco_cache_query(dim, oracle.p(flat), oracle.p(centre), oracle.p(v1), oracle.p(v2)) == 1
    e1, e2 = numpy_query(flat, dim, centre)
    assert np.allclose(v1, e1, rtol=2e-5) and np.allclose(v2, e2, rtol=2e-5)
    # only the first five in insertion order (5, 17, 90, 400, 700) take part
    e = entries.copy()
    for i in (1200, 1800, 2500, 2999):
        e[i, 1:, :] = 1e6
    w1 = np.zeros(dim, np.float32)
    w2 = np.zeros(dim, np.float32)
    assert oracle.L.lmco_cache_query(dim, oracle.p(np.ascontiguousarray(e.reshape(-1))), oracle.p(centre), oracle.p(w1), oracle.p(w2)) == 1
    assert np.array_equal(v1, w1) and np.array_equal(v2, w2)
    # nothing within the radius -> no match (the caller falls back to the isotropic Gaussian)
    far = (centre + np.float32(0.3)).astype(np.float32)
    assert oracle.L.lmco_cache_query(dim, oracle.p(flat), oracle.p(far), oracle.p(v1), oracle.p(v2)) == 0


def _cache_run(oracle, chains, steps, threads):
    h = oracle.load(os.path.join(SCENES, "torus", "lmc.xml"))
    oracle.set_option(h, "maxdepth", 6)
    oracle.set_option(h, "globalcache", 1)
    norm, ls = oracle.mlt_init(h, 100000, chains, 32)
    return oracle.run_chains(h, chains, steps, norm, ls, samples_per_chain=steps, threads=threads), (norm, ls)


def test_oracle_cache_fills_and_replaces_gradients(oracle):
    chains, steps = 4096, 120
    (f, tr, a, st), (norm, ls) = _cache_run(oracle, chains, steps, 8)
    counts = [int(x) for x in st[13:18]]
    assert counts[0] == 0                       # MLT paths have length >= 3: no D = 4 states
    assert max(counts) == 3000 and all(0 <= c <= 3000 for c in counts)
    assert int(st[11]) > 0                      # ready dimensions are queried ...
    assert int(st[12]) <= int(st[11])
    # ... and no longer differentiated: fewer gradient evaluations than the same job without the cache
    h0 = oracle.load(os.path.join(SCENES, "torus", "lmc.xml"))
    oracle.set_option(h0, "maxdepth", 6)
    f0, tr0, a0, st0 = oracle.run_chains(h0, chains, steps, norm, ls, samples_per_chain=steps)
    assert int(st[8]) < 0.9 * int(st0[8])
    assert [int(x) for x in st0[11:18]] == [0] * 7
    # before the first slot is ready the two runs are the same chains
    first_diff = int((tr == tr0).all(axis=0).argmin())
    assert first_diff > 10 and np.array_equal(tr[:, :first_diff], tr0[:, :first_diff])
    assert abs(float(f.sum()) - float(f0.sum())) < 0.02 * float(f0.sum())
    # ready slots are queried through the grid; the linear scan gives the same chains
    oracle.L.lmco_cache_use_grid(0)
    try:
        (f3, tr3, a3, st3), _ = _cache_run(oracle, chains, steps, 8)
    finally:
        oracle.L.lmco_cache_use_grid(1)
    assert np.array_equal(tr, tr3) and np.array_equal(a.view(np.uint32), a3.view(np.uint32)) and np.array_equal(st, st3)
    # the fill order is defined (chain order per iteration): independent of the number of worker threads
    (f2, tr2, a2, st2), _ = _cache_run(oracle, chains, steps, 3)
    assert np.array_equal(tr, tr2) and np.array_equal(a.view(np.uint32), a2.view(np.uint32))
    assert np.array_equal(st, st2)


@pytest.mark.gpu
def test_cuda_cache_run_bit_identical_to_oracle(lmc, oracle):
    chains, steps = 4096, 120
    (of, otr, oa, ost), (norm, ls) = _cache_run(oracle, chains, steps, 8)
    sc = lmc.ParseScene(os.path.join(SCENES, "torus", "lmc.xml"))
    sc.options.update({"maxdepth": 6, "globalcache": 1})
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, ls, samples_per_chain=steps)
    tr, a = ctx.run(steps, trace=True, a_trace=True)
    st = ctx.stats()
    film = ctx.film()
    ctx.close()
    assert st["cache_count"] == [int(x) for x in ost[13:18]] and max(st["cache_count"]) == 3000
    assert st["cache_queries"] == int(ost[11]) and st["cache_hits"] == int(ost[12])
    assert st["gradient_evals"] == int(ost[8])
    assert np.array_equal(tr, otr), "%d chains diverge" % int((tr != otr).any(axis=1).sum())
    assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert np.allclose(film, of, rtol=1e-4, atol=1e-5)
