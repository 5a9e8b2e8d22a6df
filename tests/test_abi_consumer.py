"""The drop-in boundary used from COMPILED C++ the way the reference's mlt.cpp would use it:
tests/abi_consumer.cpp includes include/lmc/{parsescene,mlt,mutation,bsdf}.h and links liblmc_b200.so.
Also: the film output functions (MergeBuffer / WriteImage) and the multi-GPU film all-reduce behind the ABI."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT, SCENES


def build_consumer(tmp_path):
    exe = str(tmp_path / "abi_consumer")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_consumer.cpp"),
           "-o", exe, "-L", PKG_DIR, "-llmc_b200", "-Wl,-rpath," + PKG_DIR]
    subprocess.check_call(cmd)
    return exe


def read_exr(path):
    """Independent minimal reader of uncompressed scan-line OpenEXR files with float channels -> H x W x 3 (RGB)."""
    b = open(path, "rb").read()
    assert b[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and b[4] == 2
    o = 8
    attrs = {}
    while b[o] != 0:
        e = b.index(0, o); name = b[o:e].decode(); o = e + 1
        e = b.index(0, o); typ = b[o:e].decode(); o = e + 1
        (size,) = struct.unpack_from("<i", b, o); o += 4
        attrs[name] = (typ, b[o:o + size]); o += size
    o += 1
    assert attrs["compression"][1] == b"\x00" and attrs["lineOrder"][1] == b"\x00"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, c, p = [], attrs["channels"][1], 0
    while c[p] != 0:
        e = c.index(0, p); chans.append(c[p:e].decode()); p = e + 1
        assert struct.unpack_from("<i", c, p)[0] == 2      # FLOAT
        p += 16
    assert chans == ["B", "G", "R"]
    offsets = struct.unpack_from("<%dQ" % h, b, o)
    img = np.zeros((h, w, 3), np.float32)
    for y in range(h):
        yy, nbytes = struct.unpack_from("<2i", b, offsets[y])
        assert yy == y + y0 and nbytes == w * 12
        row = np.frombuffer(b, np.float32, w * 3, offsets[y] + 8).reshape(3, w)
        img[y, :, 2], img[y, :, 1], img[y, :, 0] = row[0], row[1], row[2]
    return img


def test_write_image_exr_and_pfm_round_trip(lmc, tmp_path):
    rng = np.random.default_rng(7)
    film = rng.uniform(-1.0, 40.0, size=(37, 53, 3)).astype(np.float32)
    film[3, 5] = [np.float32(1e-30), np.float32(3e30), 0.0]
    p = str(tmp_path / "a.exr")
    lmc.WriteImage(p, film)
    assert np.array_equal(read_exr(p), film)
    try:
        os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
        import cv2
        img = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    except Exception:
        img = None
    if img is not None:                                   # a second, third-party reader agrees (OpenCV is BGR)
        assert np.array_equal(img[:, :, ::-1], film)
    q = str(tmp_path / "a.pfm")
    lmc.WriteImage(q, film)
    raw = open(q, "rb").read()
    head = b"PF\n53 37\n-1.0\n"
    assert raw.startswith(head)
    assert np.array_equal(np.frombuffer(raw[len(head):], np.float32).reshape(37, 53, 3)[::-1], film)


def test_merge_buffer_is_the_reference_formula(lmc):
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 5, (8, 9, 3)).astype(np.float32)
    b = rng.uniform(0, 5, (8, 9, 3)).astype(np.float32)
    out = lmc.MergeBuffer(a, 1.0 / 256, b, 1.0 / 245)
    assert np.array_equal(out, np.float32(1.0 / 256) * a + np.float32(1.0 / 245) * b)


def test_consumer_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """The compiled consumer links against the C ABI with nothing but the headers of include/lmc/.  On a machine
    without a CUDA device it must stop with the library's error, not compute anything on the host."""
    exe = build_consumer(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe, os.path.join(SCENES, "torus", "lmc.xml"), str(tmp_path / "out"), "1", "1024", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "RESULT" not in r.stdout


def parse_result(stdout):
    line = [l for l in stdout.splitlines() if l.startswith("RESULT")][0]
    return dict(kv.split("=", 1) for kv in line.split()[1:])


@pytest.mark.gpu
def test_consumer_renders_like_the_python_api(lmc, tmp_path):
    """ParseScene -> DptOptions -> MLT() from compiled C++: every mutation accounted for, the film it writes (EXR) is
    the merged direct + indirect image, and it equals the same job run through the ctypes binding."""
    exe = build_consumer(tmp_path)
    xml = os.path.join(SCENES, "torus", "lmc.xml")
    r = subprocess.run([exe, xml, str(tmp_path / "out"), "2", "8192", "1", "1"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    res = parse_result(r.stdout)
    W, H = int(res["width"]), int(res["height"])
    per_chain = 2 * W * H // 8192
    assert int(res["proposed"]) == per_chain * 8192 and int(res["mala"]) > 0 and int(res["grads"]) > 0
    assert res["finite"] == "1" and res["bsdf_size"] == "10" and int(res["intermediate"]) >= 1
    assert os.path.exists(str(tmp_path / "intermediate.exr"))
    film = read_exr(res["exr"] if os.path.isabs(res["exr"]) else str(tmp_path / res["exr"]))
    assert film.shape == (H, W, 3) and abs(float(film.mean()) - float(res["mean"])) <= 1e-5 * abs(float(res["mean"]))
    # the same job through the Python binding (same seeds, same options): same image up to fp32 atomic order
    sc = lmc.ParseScene(xml)
    sc.options.update({"maxdepth": 6})
    ctx = lmc.ChainContext(sc, 0)
    norm, init_ls = ctx.mlt_init(100000, 8192, 65536)
    assert abs(norm - float(res["norm"])) <= 1e-6 * norm
    direct = ctx.direct_lighting(4)
    ctx.begin(8192, norm, init_ls, samples_per_chain=per_chain)
    ctx.run(per_chain)
    spp = per_chain * 8192 / float(W * H)
    ref = lmc.MergeBuffer(direct, 1.0 / 4, ctx.film(), 1.0 / spp)
    ctx.close()
    assert np.allclose(film, ref, rtol=2e-3, atol=1e-4 * float(ref.max()))


@pytest.mark.gpu
def test_two_gpu_film_equals_one_gpu_film(tmp_path):
    """Chains sharded over two GPUs (lmc_create_multi, seeds = global chain ids) + the NCCL all-reduce of the film
    give the image of the one-GPU job up to fp32 summation order (SURVEY s8e determinism contract)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    exe = build_consumer(tmp_path)
    xml = os.path.join(SCENES, "torus", "lmc.xml")
    films, results = [], []
    for ndev in (1, 2):
        d = tmp_path / ("g%d" % ndev)
        d.mkdir()
        r = subprocess.run([exe, xml, str(d / "out"), "2", "8192", str(ndev)], capture_output=True, text=True, cwd=str(d))
        assert r.returncode == 0, r.stderr
        res = parse_result(r.stdout)
        results.append(res)
        films.append(read_exr(res["exr"]))
    assert results[0]["proposed"] == results[1]["proposed"] and results[0]["grads"] == results[1]["grads"]
    assert results[0]["mala"] == results[1]["mala"]
    rel = abs(float(results[0]["indirect_sum"]) - float(results[1]["indirect_sum"])) / float(results[0]["indirect_sum"])
    assert rel <= 1e-4, rel
    assert np.allclose(films[0], films[1], rtol=1e-4, atol=1e-5 * float(films[0].max()))
