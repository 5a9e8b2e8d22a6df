"""Chain-loop parity: the CUDA persistent-chain kernel vs the CPU oracle (bit-level twin).

BASELINE.json configs[0]: torus, LMC, path length 4, 1024 chains x 100 mutations, fixed PCG
seeds (chain id + seedoffset): accept/reject + step-type sequence and acceptance probabilities
must be BIT-identical; the film equal up to fp32 atomic summation order."""
import os

import numpy as np
import pytest

from conftest import SCENES


def oracle_run(oracle, xml, opts, chains, steps, norm, init_ls, **kw):
    h = oracle.load(xml)
    for k, v in opts.items():
        oracle.set_option(h, k, v)
    return oracle.run_chains(h, chains, steps, norm, init_ls, **kw)


def test_oracle_chain_loop_properties(oracle, torus_xml):
    """CPU-only sanity of the oracle itself: determinism, sharding invariance, step mix."""
    h = oracle.load(torus_xml)
    oracle.set_option(h, "maxdepth", 4)
    norm, init_ls = oracle.mlt_init(h, 40000, 256, 32)
    f1, t1, a1, s1 = oracle.run_chains(h, 256, 40, norm, init_ls, threads=4)
    f2, t2, a2, s2 = oracle.run_chains(h, 256, 40, norm, init_ls, threads=2)
    assert np.array_equal(t1, t2) and np.array_equal(a1.view(np.uint32), a2.view(np.uint32))
    # a shard [128, 256) of the same job reproduces the same chains (seed = global chain id)
    f3, t3, a3, s3 = oracle.run_chains(h, 128, 40, norm, init_ls, chain_base=128, total_chains=256, threads=2)
    assert np.array_equal(t3, t1[128:])
    # every chain starts with a large step (init states are invalid, src/mlt.h:124)
    assert ((t1[:, 0] & 3) == 0).all()
    assert int(s1[:4].sum()) == 256 * 40
    types = t1 & 3
    assert (types == 3).mean() > 0.5          # MALA small steps dominate
    assert s1[8] > 0                          # gradients were evaluated
    assert np.isfinite(f1).all() and f1.sum() > 0


def test_oracle_R_vs_T_divergence_statistics(oracle, torus_xml, ref_mala):
    """Two-tier oracle: Oracle-R = the same chain loop with the REFERENCE's own
    generated reverse-mode gradient (oracle/_ref); Oracle-T = the twin the GPU matches bit for bit.
    With the reverse sweep in the reference's merge order (option adjointcompat = 1, the default, App. B#13) the
    two gradients agree to fp32 noise, so >= 95 % of the chains must produce the IDENTICAL 100-step decision
    string (measured 99.0 %; with the true gradient it was 73.6 %); the rest are chains where an acceptance
    probability sat within rounding of its uniform draw."""
    h = oracle.load(torus_xml)
    oracle.set_option(h, "maxdepth", 4)
    chains, steps = 1024, 100            # BASELINE configs[0]
    norm, ls = oracle.mlt_init(h, 300000, chains, 32)
    fT, tT, aT, sT = oracle.run_chains(h, chains, steps, norm, ls, samples_per_chain=steps)
    assert oracle.use_reference_gradient(True) == 42
    try:
        fR, tR, aR, sR = oracle.run_chains(h, chains, steps, norm, ls, samples_per_chain=steps)
    finally:
        oracle.use_reference_gradient(False)
    same = (tT == tR).all(axis=1).mean()
    first = np.where(tT != tR, np.arange(steps)[None, :], steps).min(axis=1)
    accT, accR = sT[7] / sT[3], sR[7] / sR[3]
    print("Oracle-R vs Oracle-T: identical 100-step decision strings %.1f%% of chains; median first divergence %d; "
          "MALA acceptance T %.4f R %.4f; film sum T %.1f R %.1f" % (100 * same, int(np.median(first)), accT, accR, fT.sum(), fR.sum()))
    assert same >= 0.95
    assert abs(accT - accR) < 0.002
    assert abs(fT.sum() - fR.sum()) < 0.005 * fR.sum()
    # step-type sequences agree wherever the chains have not diverged: first steps are always equal
    assert np.array_equal(tT[:, 0], tR[:, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("maxdepth,chains,steps", [(4, 1024, 100), (8, 512, 48), (12, 256, 24)])
def test_cuda_trace_bit_identical_to_oracle(lmc, oracle, torus_xml, maxdepth, chains, steps):
    sc = lmc.ParseScene(torus_xml)
    sc.options["maxdepth"] = maxdepth
    norm, init_ls = lmc.MLTInit(sc, 300000, chains, 32)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    film = ctx.film()
    st = ctx.stats()
    ofilm, otrace, oa, ostats = oracle_run(oracle, torus_xml, {"maxdepth": maxdepth}, chains, steps, norm, init_ls,
                                           samples_per_chain=steps)
    assert np.array_equal(trace, otrace), "%d chains diverge" % int((trace != otrace).any(axis=1).sum())
    assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert st["proposed"] == [int(x) for x in ostats[:4]] and st["accepted"] == [int(x) for x in ostats[4:8]]
    assert st["gradient_evals"] == int(ostats[8]) and st["gradient_nonfinite"] == int(ostats[9])
    assert np.allclose(film, ofilm, rtol=1e-4, atol=1e-5 * max(1.0, float(ofilm.max())))
    assert abs(float(film.sum()) - float(ofilm.sum())) <= 1e-4 * float(ofilm.sum())


@pytest.mark.gpu
def test_cuda_door_scene_and_isotropic_kernel(lmc, oracle, door_xml):
    """veach-door (area light, connections, light tracing, textures) and mala=false (SmallStep)."""
    for mala in (1, 0):
        sc = lmc.ParseScene(door_xml)
        sc.options.update({"maxdepth": 6, "mala": mala})
        chains, steps = 512, 40
        norm, init_ls = lmc.MLTInit(sc, 100000, chains, 32)
        ctx = lmc.ChainContext(sc, 0)
        ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
        trace, a = ctx.run(steps, trace=True, a_trace=True)
        ofilm, otrace, oa, ostats = oracle_run(oracle, door_xml, {"maxdepth": 6, "mala": mala}, chains, steps, norm, init_ls,
                                               samples_per_chain=steps)
        assert np.array_equal(trace, otrace) and np.array_equal(a.view(np.uint32), oa.view(np.uint32))
        if not mala:
            assert ((trace & 3) != 3).all()


def compare_slices_with_oracle(oracle, xml, opts, trace, a, norm, init_ls, steps, slices):
    """trace / a: [chain][step] of the FULL job on the GPU; every (base, count) slice must equal the CPU oracle's run
    of exactly those global chain ids (seed = chain id, same init scores, same total chain count)."""
    h = oracle.load(xml)
    for k, v in opts.items():
        oracle.set_option(h, k, v)
    total = trace.shape[0]
    for base, count in slices:
        _, ot, oa, _ = oracle.run_chains(h, count, steps, norm, init_ls, chain_base=base, total_chains=total,
                                         samples_per_chain=steps)
        g, ga = trace[base:base + count], a[base:base + count]
        assert np.array_equal(g, ot), "chains %d..%d: %d diverge" % (base, base + count, int((g != ot).any(axis=1).sum()))
        assert np.array_equal(ga.view(np.uint32), oa.view(np.uint32))


@pytest.mark.gpu
def test_cuda_launch_split_and_sharding_invariance(lmc, torus_xml):
    """100 mutations in one launch == 4 launches of 25 (state round-trips through HBM bit-exactly),
    and a shard of the chain range reproduces the same chains."""
    sc = lmc.ParseScene(torus_xml)
    sc.options["maxdepth"] = 8
    chains, steps = 512, 100
    norm, init_ls = lmc.MLTInit(sc, 300000, chains, 32)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    t_one, a_one = ctx.run(steps, trace=True, a_trace=True)
    film_one = ctx.film()
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    parts = [ctx.run(25, trace=True, a_trace=True) for _ in range(4)]
    t_split = np.concatenate([p[0] for p in parts], axis=1)
    assert np.array_equal(t_one, t_split)
    assert np.allclose(ctx.film(), film_one, rtol=1e-4, atol=1e-5)
    ctx.begin(128, norm, init_ls, chain_base=256, total_chains=chains, samples_per_chain=steps)
    t_shard, _ = ctx.run(steps, trace=True)
    assert np.array_equal(t_shard, t_one[256:384])


@pytest.mark.gpu
def test_cuda_full_size_properties(lmc, oracle, torus_xml):
    """BASELINE configs[1] size (2^20 chains, maxdepth 8): size-independent invariants."""
    sc = lmc.ParseScene(torus_xml)
    sc.options["maxdepth"] = 8
    chains, steps = 1 << 20, 8
    norm, init_small = lmc.MLTInit(sc, 300000, 4096, 32)
    init_ls = np.resize(init_small, chains)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    st = ctx.stats()
    assert sum(st["proposed"]) == chains * steps
    assert st["proposed"][0] >= chains                 # every chain opens with a large step
    assert all(a <= p for a, p in zip(st["accepted"], st["proposed"]))
    film = ctx.film()
    assert np.isfinite(film).all() and film.min() >= 0.0
    # slices of the 2^20-chain job against the CPU oracle, bit for bit (decisions and acceptance probabilities)
    compare_slices_with_oracle(oracle, torus_xml, {"maxdepth": 8}, trace, a, norm, init_ls, steps,
                               [(0, 512), (777777, 256), (chains - 256, 256)])
    # energy check: sum of splats / mutations estimates the image mean (normalization = b)
    mean = float(film.sum()) / (chains * steps) / 3.0
    assert 0.2 * norm < mean < 5.0 * norm
    # the first 1024 chains of the big job are the chains of a 1024-chain job (seed = chain id)
    ctx.begin(1024, norm, init_ls, total_chains=chains, samples_per_chain=steps)
    t_small, _ = ctx.run(steps, trace=True)
    assert ((t_small[:, 0] & 3) == 0).all()


@pytest.mark.gpu
def test_cuda_per_vertex_wavefront_equals_monolithic_propose(lmc, torus_xml, door_xml, monkeypatch):
    """The two device forms of the proposal phase -- per-vertex wavefront (k_prop_start / k_trace /
    k_shade<stage> / k_shade_tail / k_shadow / k_prop_post, the default) and the monolithic
    k_wave_propose kept as an A/B switch (LMC_WAVEFRONT=0) -- give the same chains bit for bit."""
    for xml, depth in ((torus_xml, 8), (door_xml, 10)):
        out = []
        for wf in ("1", "0"):
            monkeypatch.setenv("LMC_WAVEFRONT", wf)
            sc = lmc.ParseScene(xml)
            sc.options["maxdepth"] = depth
            chains, steps = 2048, 24
            norm, init_ls = lmc.MLTInit(sc, 100000, chains, 32)
            ctx = lmc.ChainContext(sc, 0)
            ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
            trace, a = ctx.run(steps, trace=True, a_trace=True)
            out.append((trace, a, ctx.film(), ctx.stats()))
            ctx.close()
        (t1, a1, f1, s1), (t0, a0, f0, s0) = out
        assert np.array_equal(t1, t0) and np.array_equal(a1.view(np.uint32), a0.view(np.uint32))
        assert s1["proposed"] == s0["proposed"] and s1["accepted"] == s0["accepted"]
        assert s1["gradient_evals"] == s0["gradient_evals"]
        assert s1["kernel_launches"] > s0["kernel_launches"]      # really two different launch sequences
        assert np.allclose(f1, f0, rtol=1e-4, atol=1e-5 * max(1.0, float(f0.max())))


@pytest.mark.gpu
@pytest.mark.parametrize("scene,opts,steps", [
    ("door", {"maxdepth": 12}, 4),                          # BASELINE configs[2]: veach-door, LMC, path length 12, 2^20 chains
    ("torus_h2mc", {"maxdepth": 8}, 3),                     # BASELINE configs[3]: torus, H2MC mutation, path length 8, 2^20 chains
])
def test_cuda_full_size_other_configs(lmc, oracle, torus_xml, door_xml, scene, opts, steps):
    """The other single-GPU BASELINE configurations at their full chain count: size-independent invariants
    (every mutation accounted for, first step large, finite non-negative film with the right energy) and the
    first 512 chains of the big job equal to the chains of a 512-chain job (seed = global chain id)."""
    import os
    xml = door_xml if scene == "door" else os.path.join(os.path.dirname(torus_xml), "h2mc.xml")
    sc = lmc.ParseScene(xml)
    sc.options.update(opts)
    chains = 1 << 20
    ctx = lmc.ChainContext(sc, 0)
    norm, init_small = ctx.mlt_init(300000, 4096, 4096)
    init_ls = np.resize(init_small, chains)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    st = ctx.stats()
    assert sum(st["proposed"]) == chains * steps
    assert st["proposed"][0] >= chains
    assert all(a <= p for a, p in zip(st["accepted"], st["proposed"]))
    # slices of the full-size job against the CPU oracle, bit for bit
    compare_slices_with_oracle(oracle, xml, opts, trace, a, norm, init_ls, steps,
                               [(0, 256), (500000, 128)] if scene == "door" else [(0, 96), (900001, 64)])
    if scene == "torus_h2mc":
        assert st["proposed"][2] > 0 and st["proposed"][3] == 0      # H2MC small steps, no MALA
    else:
        assert st["proposed"][3] > 0
    assert st["gradient_evals"] > 0
    film = ctx.film()
    # Negative splats exist in the reference's arithmetic too (the env map's bilinear lookup extrapolates
    # across the wrap-around column); the CPU oracle reproduces them bit for bit
    # (tools/neg_hunt.py).  They are rare: bound their mass instead of forbidding them.
    assert np.isfinite(film).all() and float(film[film < 0].sum()) > -1e-4 * float(film.sum())
    mean = float(film.sum()) / (chains * steps) / 3.0
    assert 0.2 * norm < mean < 5.0 * norm
    ctx.begin(512, norm, init_ls, total_chains=chains, samples_per_chain=steps)
    t_small, a_small = ctx.run(steps, trace=True, a_trace=True)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    ctx.run(steps)                                  # same big job again: deterministic film up to atomic order
    film2 = ctx.film()
    assert np.allclose(film, film2, rtol=1e-3, atol=1e-4 * max(1.0, float(film.max())))
    assert ((t_small[:, 0] & 3) == 0).all()
    ctx.close()


@pytest.mark.gpu
def test_cuda_proposal_path_is_chosen_by_chain_count(lmc, torus_xml, monkeypatch):
    """Without LMC_WAVEFRONT the library runs small jobs through the monolithic proposal kernel (about 11 launches
    per iteration) and large ones through the per-vertex wavefront (about 50); both agree with a forced run."""
    sc = lmc.ParseScene(torus_xml)
    sc.options["maxdepth"] = 8
    chains, steps = 4096, 8
    norm, init_ls = lmc.MLTInit(sc, 100000, chains, 32)
    runs = {}
    for mode in ("auto", "1"):
        if mode == "auto":
            monkeypatch.delenv("LMC_WAVEFRONT", raising=False)
        else:
            monkeypatch.setenv("LMC_WAVEFRONT", mode)
        ctx = lmc.ChainContext(sc, 0)
        ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
        l0 = ctx.stats()["kernel_launches"]
        trace, _ = ctx.run(steps, trace=True)
        runs[mode] = (trace, (ctx.stats()["kernel_launches"] - l0) / steps)
        ctx.close()
    assert np.array_equal(runs["auto"][0], runs["1"][0])
    assert runs["auto"][1] < 15 < runs["1"][1]


def _outlier_opts():
    # thresholds of src/mutation.h:5-8 lowered so that the reset of src/mlt.cpp:147-169 fires within a short run:
    # any chain rejected 6 times in a row is reset (weak rule), a chain whose current lsScore exceeds 0.5 x
    # normalization already after 2 (strong rule), and the walk over the init states skips those above 0.5 x too
    return {"maxdepth": 6, "outlierweakrejectcnt": 6, "outlierstrongrejectcnt": 2, "outlierratiothreshold": 0.5}


def test_oracle_outlier_reset_fires_and_follows_the_reference_rule(oracle, torus_xml):
    """Outlier ("stuck chain") reset, src/mlt.cpp:147-169 with the constants of src/mutation.h:5-8 as options.
    With the default thresholds (10000 / 1000 / 30) no 100-step test ever reaches the branch; lowered thresholds
    make it fire.  Checked against the rule itself: after a reset the chain is invalid, so the next step is a large
    step whatever the uniform draw says, and the reset count of the run is reproduced by a replay of the decision
    strings."""
    h = oracle.load(torus_xml)
    opts = _outlier_opts()
    for k, v in opts.items():
        oracle.set_option(h, k, v)
    chains, steps = 512, 160
    norm, ls = oracle.mlt_init(h, 100000, chains, 32)
    film, trace, a, stats = oracle.run_chains(h, chains, steps, norm, ls, samples_per_chain=steps)
    resets = int(stats[10])
    assert resets > 50, "the lowered thresholds must trigger resets (%d)" % resets
    # replay: adjacentReject counts consecutive rejections (never cleared by a reset, as in the reference);
    # the weak rule alone gives a lower bound on the resets, and every reset forces a large step next
    acc = (trace >> 2) & 1
    typ = trace & 3
    lower = 0
    for c in range(chains):
        adj = 0
        for s in range(steps):
            if acc[c, s]:
                adj = 0
            else:
                adj += 1
                if adj > opts["outlierweakrejectcnt"]:
                    lower += 1
                    if s + 1 < steps:
                        assert typ[c, s + 1] == 0, "chain %d step %d: a reset chain must restart with a large step" % (c, s + 1)
    assert lower <= resets
    assert np.isfinite(film).all()
    # default thresholds: the same run never resets
    h2 = oracle.load(torus_xml)
    oracle.set_option(h2, "maxdepth", 6)
    _, _, _, stats2 = oracle.run_chains(h2, chains, steps, norm, ls, samples_per_chain=steps)
    assert int(stats2[10]) == 0


@pytest.mark.gpu
def test_cuda_outlier_reset_bit_identical_to_oracle(lmc, oracle, torus_xml):
    """The reset branch on the device (k_wave_finish) against the CPU oracle: same decisions, same acceptance
    probabilities, same number of resets, same film."""
    opts = _outlier_opts()
    sc = lmc.ParseScene(torus_xml)
    sc.options.update(opts)
    chains, steps = 512, 160
    norm, init_ls = lmc.MLTInit(sc, 100000, chains, 32)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    st = ctx.stats()
    film = ctx.film()
    ofilm, otrace, oa, ostats = oracle_run(oracle, torus_xml, opts, chains, steps, norm, init_ls, samples_per_chain=steps)
    assert st["outlier_resets"] == int(ostats[10]) and st["outlier_resets"] > 50
    assert np.array_equal(trace, otrace)
    assert np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert np.allclose(film, ofilm, rtol=1e-4, atol=1e-5)
    ctx.close()


@pytest.mark.gpu
def test_cuda_textured_parameter_scene_bit_identical(lmc, oracle):
    """scenes/torus/textured.xml: bitmap-textured Phong Ks / exponent and RoughDielectric alpha / Kt:
    the device evaluates the parameter maps per hit exactly like the host twin."""
    xml = os.path.join(SCENES, "torus", "textured.xml")
    sc = lmc.ParseScene(xml)
    sc.options["maxdepth"] = 6
    chains, steps = 512, 40
    norm, init_ls = lmc.MLTInit(sc, 100000, chains, 32)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    film = ctx.film()
    ofilm, otrace, oa, ostats = oracle_run(oracle, xml, {"maxdepth": 6}, chains, steps, norm, init_ls, samples_per_chain=steps)
    assert np.array_equal(trace, otrace) and np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert np.allclose(film, ofilm, rtol=1e-4, atol=1e-5 * max(1.0, float(ofilm.max())))


@pytest.mark.gpu
@pytest.mark.parametrize("chains,opts", [
    (1, {"maxdepth": 8}), (33, {"maxdepth": 8}), (1000, {"maxdepth": 3}), (257, {"maxdepth": 8, "h2mc": 1, "mala": 0}),
    (1000, {"maxdepth": 6, "globalcache": 1}), (77, {"maxdepth": 12}),
])
def test_cuda_ragged_sizes_bit_identical(lmc, oracle, torus_xml, chains, opts):
    """Chain counts that are no multiple of a warp / block / list alignment, the shallowest depths, every mutation kind:
    grids, list padding and queue sizing must not depend on round numbers (wavefront path forced by conftest)."""
    sc = lmc.ParseScene(torus_xml)
    sc.options.update(opts)
    steps = 24
    norm, init_ls = lmc.MLTInit(sc, 30000, chains, 32)
    ctx = lmc.ChainContext(sc, 0)
    ctx.begin(chains, norm, init_ls, samples_per_chain=steps)
    trace, a = ctx.run(steps, trace=True, a_trace=True)
    st = ctx.stats()
    ctx.close()
    ofilm, otrace, oa, ostats = oracle_run(oracle, torus_xml, opts, chains, steps, norm, init_ls, samples_per_chain=steps)
    assert np.array_equal(trace, otrace) and np.array_equal(a.view(np.uint32), oa.view(np.uint32))
    assert st["proposed"] == [int(x) for x in ostats[:4]] and st["gradient_evals"] == int(ostats[8])
