"""Host loader + BVH2 (the Embree replacement): traversal must return exactly what a brute-force
closest-hit with the same Moeller-Trumbore leaf test returns (oracle-defined semantics,
SURVEY.md s8c: Embree's own tie-breaking is unpinned)."""
import ctypes

import numpy as np
import pytest


def make_rays(n, seed, center, radius):
    rng = np.random.default_rng(seed)
    org = center + rng.normal(size=(n, 3)) * radius * 0.6
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([org, d], axis=1).astype(np.float32)


def scene_center_radius(oracle, h):
    s = oracle.scene_serialized(h)
    return s[34:37].astype(np.float64), float(s[37]) / 1000.0


@pytest.mark.parametrize("scene", ["torus", "veachdoor"])
def test_bvh_equals_brute_force(oracle, torus_xml, door_xml, scene):
    h = oracle.load(torus_xml if scene == "torus" else door_xml)
    c, r = scene_center_radius(oracle, h)
    rays = make_rays(3000, 5, c, r)
    out = {}
    for brute in (0, 1):
        tid = np.zeros(len(rays), np.int32)
        tuv = np.zeros((len(rays), 3), np.float32)
        oracle.L.lmco_intersect(h, len(rays), oracle.p(rays), ctypes.c_float(5e-4), ctypes.c_float(np.inf), brute,
                                oracle.p(tid), oracle.p(tuv))
        out[brute] = (tid, tuv)
    assert (out[0][0] >= 0).mean() > 0.1           # the probe actually hits geometry
    assert np.array_equal(out[0][0], out[1][0])
    hit = out[0][0] >= 0
    assert np.array_equal(out[0][1][hit].view(np.uint32), out[1][1][hit].view(np.uint32))


def test_empty_and_degenerate_queries(oracle, torus_xml):
    h = oracle.load(torus_xml)
    tid = np.zeros(1, np.int32)
    tuv = np.zeros((1, 3), np.float32)
    # zero rays: nothing written, no crash
    assert oracle.L.lmco_intersect(h, 0, None, ctypes.c_float(0), ctypes.c_float(1), 0, oracle.p(tid), oracle.p(tuv)) == 0
    # a ray pointing away from everything misses
    ray = np.array([[0, 0, 1e6, 0, 0, 1]], np.float32)
    oracle.L.lmco_intersect(h, 1, oracle.p(ray), ctypes.c_float(5e-4), ctypes.c_float(np.inf), 0, oracle.p(tid), oracle.p(tuv))
    assert tid[0] == -1
    # an empty [tmin, tmax] interval misses
    c, r = scene_center_radius(oracle, h)
    rays = make_rays(200, 9, c, r)
    tids = np.zeros(200, np.int32)
    tuvs = np.zeros((200, 3), np.float32)
    oracle.L.lmco_intersect(h, 200, oracle.p(rays), ctypes.c_float(1.0), ctypes.c_float(0.5), 0, oracle.p(tids), oracle.p(tuvs))
    assert (tids == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["torus", "veachdoor"])
def test_cuda_bvh_probe_matches_oracle(lmc, oracle, torus_xml, door_xml, scene):
    xml = torus_xml if scene == "torus" else door_xml
    h = oracle.load(xml)
    c, r = scene_center_radius(oracle, h)
    rays = make_rays(20000, 11, c, r)
    tid_o = np.zeros(len(rays), np.int32)
    tuv_o = np.zeros((len(rays), 3), np.float32)
    oracle.L.lmco_intersect(h, len(rays), oracle.p(rays), ctypes.c_float(5e-4), ctypes.c_float(np.inf), 0, oracle.p(tid_o), oracle.p(tuv_o))
    ctx = lmc.ChainContext(lmc.ParseScene(xml), 0)
    tid, gp, tuv = ctx.bvh_probe(rays, 5e-4, np.inf)
    assert np.array_equal(tid, tid_o)
    hit = tid >= 0
    assert np.array_equal(tuv[hit].view(np.uint32), tuv_o[hit].view(np.uint32))
    occ, _, _ = ctx.bvh_probe(rays, 5e-4, np.inf, any_hit=True)
    assert np.array_equal(occ.astype(bool), hit)
    empty_tid, _, _ = ctx.bvh_probe(np.zeros((0, 6), np.float32), 0.0, 1.0)
    assert empty_tid.shape == (0,)
