"""Host loader + BVH2 (the Embree replacement): traversal must return exactly what a brute-force
closest-hit with the same Moeller-Trumbore leaf test returns (oracle-defined semantics,
SURVEY.md s8c: Embree's own tie-breaking is unpinned)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES


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


def _read_rawf(path):
    with open(path, "rb") as f:
        assert f.read(4) == b"RAWF"
        w, h, is8 = np.frombuffer(f.read(12), "<i4")
        if is8:
            rgb = np.frombuffer(f.read(), np.uint8).astype(np.float32) / np.float32(255.0)
        else:
            rgb = np.frombuffer(f.read(), "<f4")
    return int(w), int(h), int(is8), rgb


IMAGES = [("torus", "checker.png"), ("torus", "sunsky.exr"), ("veachdoor", "72cf.jpg"), ("veachdoor", "72rdf.jpg"),
          ("veachdoor", "checker.jpg"), ("veachdoor", "marble.jpg"), ("veachdoor", "perlin.jpg"), ("veachdoor", "pic.jpg")]


@pytest.mark.parametrize("scene,name", IMAGES)
def test_native_image_decoders_match_opencv_decode(oracle, scene, name):
    """The loader's own PNG / OpenEXR / JPEG decoders (csrc/host/image_decode.h, jpeg_decode.h -- what the reference gets
    from OpenImageIO, src/image.cpp:5-45, src/bitmaptexture.h:73-146) against the OpenCV decode of the same files kept in
    tests/golden/decoded/<file>.rawf (written by tools/stage_scenes.py): every image of the two bundled scenes, bit for
    bit -- incl. the JPEGs (four baseline 4:2:0 files with odd sizes, two progressive 4:4:4 ones)."""
    import ctypes
    path = os.path.join(SCENES, scene, "data", name)
    w, h, is8, ref = _read_rawf(os.path.join(GOLDEN, "decoded", name + ".rawf"))
    whi = np.zeros(3, np.int32)
    out = np.zeros(w * h * 3, np.float32)
    oracle.L.lmco_decode_image.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
    assert oracle.L.lmco_decode_image(path.encode(), oracle.p(whi), oracle.p(out), out.size) == 0, oracle.L.lmco_last_error()
    assert (int(whi[0]), int(whi[1]), int(whi[2])) == (w, h, is8)
    assert np.array_equal(out.view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))


def test_scene_loads_from_predecoded_containers(oracle, tmp_path):
    """The .rawf hand-over path (formats the loader does not decode): a copy of the torus scene whose images exist ONLY
    as .rawf containers loads to the same scene and runs the same first mutations as the natively decoded one."""
    import shutil
    dst = tmp_path / "torus"
    shutil.copytree(os.path.join(SCENES, "torus"), dst, ignore=shutil.ignore_patterns("*.png", "*.exr", "textured.xml"))
    for name in ("checker.png", "sunsky.exr"):
        shutil.copyfile(os.path.join(GOLDEN, "decoded", name + ".rawf"), dst / "data" / (name + ".rawf"))
    h0 = oracle.load(os.path.join(SCENES, "torus", "lmc.xml"))
    h1 = oracle.load(str(dst / "lmc.xml"))
    for h in (h0, h1):
        oracle.set_option(h, "maxdepth", 4)
    n0, l0 = oracle.mlt_init(h0, 20000, 256, 8)
    n1, l1 = oracle.mlt_init(h1, 20000, 256, 8)
    assert n0 == n1 and np.array_equal(l0, l1)
    r0 = oracle.run_chains(h0, 256, 12, n0, l0, samples_per_chain=12)
    r1 = oracle.run_chains(h1, 256, 12, n1, l1, samples_per_chain=12)
    assert np.array_equal(r0[1], r1[1]) and np.array_equal(r0[0], r1[0])


def _bsdf_records(c, l, v):
    """The 10-float BSDF records along a serialized path (layout: SURVEY.md App. A.4)."""
    o, out = 3, []
    if l > 1:
        o += 1 + 56
        for k in range(l - 1):
            out.append(v[o + 48:o + 58])
            o += 46 + 2 + 10 + (0 if k == l - 2 else 1)
    for k in range(c - 1):
        if k == c - 2:
            if l == 1:
                out.append(v[o + 46 + 56:o + 46 + 66])
            elif l >= 2:
                out.append(v[o + 46:o + 56])
        else:
            out.append(v[o + 48:o + 58])
            o += 59
    return out


def test_textured_bsdf_parameters(oracle):
    """Parse3DMap / Parse1DMap of ParseBSDF (src/parsescene.cpp:341-412): Phong specularReflectance / exponent and
    RoughDielectric alpha / specularTransmittance read from bitmaps (scenes/torus/textured.xml = lmc.xml with those four
    parameters textured).  The serialized BSDF records (BSDF::Serialize evaluates the textures at the hit point,
    src/phong.cpp:14-20, src/roughdielectric.cpp:13-20) must carry per-hit values; the constant scene must not."""
    def records(xml):
        h = oracle.load(os.path.join(SCENES, "torus", xml))
        oracle.set_option(h, "maxdepth", 6)
        rec = oracle.sample_paths(h, 7, 2500, perturb=False, max_len=6)
        phong, glass = [], []
        for r in rec:
            c, l = int(r[0]), int(r[1])
            for b in _bsdf_records(c, l, r[oracle.REC_HEAD + 25:]):
                (phong if int(b[0]) == 1 else glass if int(b[0]) == 2 else []).append(b)
        return np.array(phong), np.array(glass)
    p0, g0 = records("lmc.xml")
    p1, g1 = records("textured.xml")
    assert len(p1) > 200 and len(g1) > 200
    # constant scene: two Phong materials, one glass
    assert len(np.unique(p0[:, 7])) == 2 and len(np.unique(g0[:, 9])) == 1 and np.allclose(g0[:, 4:7], 1.0)
    # textured scene.  Phong record: type, Kd(3), Ks(3), exponent, KsWeight; glass: type, Ks(3), Kt(3), eta, invEta, alpha
    floor = p1[p1[:, 7] <= 1.0]                               # the 8-bit exponent map decodes to (0, 1] (gamma 2.2)
    assert len(floor) > 100 and len(np.unique(floor[:, 7])) > 20 and floor[:, 7].min() > 0.0
    assert len(np.unique(floor[:, 4])) >= 2 and floor[:, 4:7].min() >= 0.0 and floor[:, 4:7].max() <= 1.001   # fastpow(1, 2.2) = 1.0000076
    assert np.array_equal(floor[:, 4], floor[:, 5])           # grey checker: Ks channels equal
    alpha = g1[:, 9]
    lo, hi = (40 / 255.0) ** 2.2, (140 / 255.0) ** 2.2        # range of the alpha map's five grey levels
    assert len(np.unique(alpha)) > 20 and alpha.min() >= lo * 0.9 and alpha.max() <= hi * 1.1
    assert (alpha < 0.05).any() and (alpha > 0.05).any()      # both parametrisations of the glass vertex occur
    assert g1[:, 4:7].min() >= 0.0 and g1[:, 4:7].max() <= 1.001 and len(np.unique(g1[:, 4])) >= 2
    # KsWeight uses the texture AVERAGES (src/phong.cpp:159-169): one value per material, not per hit
    assert len(np.unique(floor[:, 8])) == 1
