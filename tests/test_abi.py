"""The C-ABI shared library: it loads without a GPU, exports every symbol include/lmc/lmc_abi.h
declares, mirrors the reference's loader / option surface, and fails loudly (no CPU fallback)
when compute is requested without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SCENES


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lmc", "lmc_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lmc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lmc):
    lib = ctypes.CDLL(lmc.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), "liblmc_b200.so does not export %s" % s
    assert b"sm_100a" in lmc.load_library().lmc_version()


def test_parse_scene_mirrors_reference_loader(lmc, torus_xml, door_xml):
    sc = lmc.ParseScene(torus_xml)
    # scenes/torus/lmc.xml: 1024x768 film, 5 serialized shapes (23 614 triangles), one envmap,
    # <dpt> spp 245, largestepprob 0.05, largestepscale 4, mala true (src/parsescene.cpp:535-590)
    assert (sc.width, sc.height) == (1024, 768)
    assert sc.info["num_triangles"] == 23614 and sc.info["num_shapes"] == 5 and sc.info["num_lights"] == 1
    assert sc.info["spp"] == 245 and sc.info["num_init_samples"] == 300000
    assert sc.options["mala"] == 1 and sc.options["h2mc"] == 0 and sc.options["maxdepth"] == 8
    assert abs(sc.options["largestepprob"] - 0.05) < 1e-7 and sc.options["largestepscale"] == 4
    # defaults of src/dptoptions.h and the compile-time knobs
    assert abs(sc.options["perturbstddev"] - 0.01) < 1e-7 and abs(sc.options["mala-stepsize"] - 0.005) < 1e-7
    assert sc.options["mala-gn"] == 100 and sc.options["pssmaxlength"] == 12 and sc.options["maxdervdepth"] == 8
    sc.options["maxdepth"] = 4
    assert sc.options["maxdepth"] == 4
    with pytest.raises(lmc.LmcError):
        sc.options["no-such-option"] = 1
    ser = sc.serialized()   # Serialize(scene): 38 floats
    assert ser.shape == (38,) and ser[32] == 1024 * 768 and ser[37] > 1000.0
    door = lmc.ParseScene(door_xml)
    assert (door.width, door.height) == (1280, 720) and door.info["num_shapes"] == 22 and door.info["num_textures"] == 6


def test_errors_are_reported_not_thrown(lmc, tmp_path):
    with pytest.raises(lmc.LmcError, match="cannot open"):
        lmc.ParseScene(str(tmp_path / "missing.xml"))
    bad = tmp_path / "bad.xml"
    bad.write_text("<scene><shape type='obj'><string name='filename' value='nope.obj'/></shape></scene>")
    with pytest.raises(lmc.LmcError):
        lmc.ParseScene(str(bad))


def test_scene_pack_round_trip(lmc, torus_xml, tmp_path):
    sc = lmc.ParseScene(torus_xml)
    p = str(tmp_path / "torus.pack")
    sc.save_pack(p)
    sc2 = lmc.ParseScene(p)
    assert sc2.info == sc.info
    assert np.array_equal(sc2.serialized(), sc.serialized())


def test_mlt_init_is_deterministic_and_matches_oracle(lmc, oracle, torus_xml):
    sc = lmc.ParseScene(torus_xml)
    sc.options["maxdepth"] = 4
    n1, l1 = lmc.MLTInit(sc, 40000, 256, 32)
    n2, l2 = lmc.MLTInit(sc, 40000, 256, 32)
    assert n1 == n2 and np.array_equal(l1, l2) and n1 > 0
    h = oracle.load(torus_xml)
    oracle.set_option(h, "maxdepth", 4)
    n3, l3 = oracle.mlt_init(h, 40000, 256, 32)
    assert n3 == n1 and np.array_equal(l3, l1)
    with pytest.raises(lmc.LmcError, match="MLT initialization failed"):
        lmc.MLTInit(sc, 100, 4096, 4)


def test_no_cpu_fallback(lmc, torus_xml):
    """Without a device lmc_create must fail with a CUDA error, never compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    sc = lmc.ParseScene(torus_xml)
    with pytest.raises(lmc.LmcError, match="no CUDA device|CUDA"):
        lmc.ChainContext(sc, 0)


@pytest.mark.gpu
def test_cuda_mlt_init_device_equals_host(lmc, torus_xml, door_xml):
    """lmc_mlt_init_device (init paths generated on the GPU, one CUDA thread per logical thread) against
    lmc_mlt_init (host restatement of MLTInit, src/mlt.h:41-154) with the same logical thread count:
    normalisation and per-chain init scores bit-equal."""
    import numpy as np
    for xml, depth, samples, chains, threads in ((torus_xml, 8, 60000, 4096, 1024), (door_xml, 6, 30000, 1000, 777),
                                                  (torus_xml, 4, 5000, 256, 32)):
        sc = lmc.ParseScene(xml)
        sc.options["maxdepth"] = depth
        norm_h, ls_h = lmc.MLTInit(sc, samples, chains, threads)
        ctx = lmc.ChainContext(sc, 0)
        norm_d, ls_d = ctx.mlt_init(samples, chains, threads)
        ctx.close()
        assert np.float32(norm_h).tobytes() == np.float32(norm_d).tobytes()
        assert np.array_equal(ls_h.view(np.uint32), ls_d.view(np.uint32))
        assert norm_d > 0 and (ls_d > 0).all()
        # the multi-GPU split: three uneven shards of the logical threads, concatenated in thread order, + the
        # sequential tail (lmc_mlt_init_device_part / lmc_mlt_init_finish) give the same bits again
        ctx = lmc.ChainContext(sc, 0)
        cuts = [0, threads // 3, threads // 3 + 1, threads]
        parts = [ctx.mlt_init_part(samples, threads, cuts[k], cuts[k + 1]) for k in range(3)]
        ctx.close()
        norm_s, ls_s = lmc.mlt_init_finish(np.concatenate(parts), samples, chains)
        assert np.float32(norm_s).tobytes() == np.float32(norm_d).tobytes()
        assert np.array_equal(ls_s.view(np.uint32), ls_d.view(np.uint32))
