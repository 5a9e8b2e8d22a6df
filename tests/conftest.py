import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "langevin-mcmc_b200")
SCENES = os.path.join(ROOT, "scenes")
GOLDEN = os.path.join(ROOT, "tests", "golden")


# The library picks the device form of the proposal phase by chain count (per-vertex wavefront from ~4e5
# chains, monolithic kernel below); the parity tests run small jobs, so they force the wavefront -- the path
# the full-size configurations use.  test_cuda_per_vertex_wavefront_equals_monolithic_propose covers the other
# form and test_cuda_proposal_path_is_chosen_by_chain_count the automatic choice.
os.environ.setdefault("LMC_WAVEFRONT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    """The package directory has a hyphen in its name: import it under the alias lmc_b200."""
    if "lmc_b200" in sys.modules:
        return sys.modules["lmc_b200"]
    spec = importlib.util.spec_from_file_location("lmc_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lmc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


class Oracle:
    """ctypes view of oracle/liblmc_oracle.so (test infrastructure; see oracle/oracle_api.cpp)."""
    VSTRIDE = 1005
    REC_HEAD = 16

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "liblmc_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ROOT, "oracle"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        self.L = ctypes.CDLL(path)
        self.L.lmco_scene_load.restype = ctypes.c_void_p
        self.L.lmco_last_error.restype = ctypes.c_char_p

    @staticmethod
    def p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def load(self, path):
        h = self.L.lmco_scene_load(path.encode())
        if not h:
            raise RuntimeError(self.L.lmco_last_error().decode())
        return ctypes.c_void_p(h)

    def set_option(self, h, name, value):
        assert self.L.lmco_set_option(h, name.encode(), ctypes.c_double(value)) == 0, name

    def info(self, h):
        a = (ctypes.c_int * 8)()
        self.L.lmco_scene_info(h, a)
        return dict(zip(["width", "height", "num_triangles", "num_bvh_nodes", "num_lights", "num_shapes", "spp",
                         "num_init_samples"], list(a)))

    def scene_serialized(self, h):
        s = np.zeros(38, np.float32)
        self.L.lmco_scene_serialized(h, self.p(s))
        return s

    def get_option(self, h, name):
        v = ctypes.c_double()
        assert self.L.lmco_get_option(h, name.encode(), ctypes.byref(v)) == 0, name
        return v.value

    def mlt_init(self, h, num_init, num_chains, logical_threads=32):
        norm = ctypes.c_float()
        ls = np.zeros(num_chains, np.float32)
        rc = self.L.lmco_mlt_init(h, ctypes.c_longlong(num_init), num_chains, logical_threads, ctypes.byref(norm), self.p(ls))
        assert rc == 0, self.L.lmco_last_error()
        return norm.value, ls

    def run_chains(self, h, num_chains, steps, norm, init_ls, chain_base=0, total_chains=None, samples_per_chain=None,
                   threads=8, want_trace=True):
        total = total_chains if total_chains is not None else num_chains
        info = self.info(h)
        film = np.zeros((info["height"], info["width"], 3), np.float32)
        trace = np.zeros((num_chains, steps), np.uint8) if want_trace else None
        a = np.zeros((num_chains, steps), np.float32) if want_trace else None
        stats = np.zeros(18, np.uint64)
        spc = samples_per_chain if samples_per_chain is not None else steps
        rc = self.L.lmco_run_chains(h, num_chains, chain_base, total, ctypes.c_longlong(steps), ctypes.c_longlong(spc),
                                    ctypes.c_float(norm), self.p(init_ls), self.p(film), self.p(trace), self.p(a),
                                    threads, self.p(stats))
        assert rc == 0, self.L.lmco_last_error()
        return film, trace, a, stats

    def use_reference_gradient(self, enable):
        """Oracle-R: route MALA gradients through oracle/_ref/libpathref_mala.so (reference's own code)."""
        path = os.path.join(ROOT, "oracle", "_ref", "libpathref_mala.so")
        if enable and not os.path.exists(path):
            return -1
        return self.L.lmco_use_reference_gradient(path.encode(), 1 if enable else 0)

    def direct_lighting(self, h, direct_spp, threads=8):
        """DirectLighting(scene, buffer), src/direct.cpp:4-54 -> unweighted sample buffer."""
        info = self.info(h)
        film = np.zeros((info["height"], info["width"], 3), np.float32)
        assert self.L.lmco_direct_lighting(h, int(direct_spp), self.p(film), threads) == 0, self.L.lmco_last_error()
        return film

    def use_staged(self, enable):
        """Run the proposal phase through the staged (per-vertex wavefront) path functions."""
        return self.L.lmco_use_staged(1 if enable else 0)

    def sample_paths(self, h, seed, num_large_steps, perturb=True, max_len=8, max_records=20000):
        rec = self.REC_HEAD + 25 + self.VSTRIDE
        out = np.zeros((max_records, rec), np.float32)
        n = self.L.lmco_sample_paths(h, ctypes.c_ulonglong(seed), num_large_steps, 1 if perturb else 0, max_len,
                                     self.VSTRIDE, max_records, self.p(out))
        return out[:n]

    def eval_batch(self, h, c, l, primary, vert, want_grad=True):
        primary = np.ascontiguousarray(primary, np.float32)
        vert = np.ascontiguousarray(vert, np.float32)
        n = primary.shape[0]
        dim = 2 * max(c + l - 1, 2)
        ll = np.zeros(n, np.float32)
        g = np.zeros((n, dim), np.float32) if want_grad else None
        self.L.lmco_eval_batch(h, c, l, n, self.p(primary), primary.shape[1], self.p(vert), vert.shape[1], self.p(ll), self.p(g))
        return ll, g


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def lmc():
    return load_package()


@pytest.fixture(scope="session")
def torus_xml():
    return os.path.join(SCENES, "torus", "lmc.xml")


@pytest.fixture(scope="session")
def door_xml():
    return os.path.join(SCENES, "veachdoor", "lmc.xml")


def ref_lib(name):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return ctypes.CDLL(p) if os.path.exists(p) else None


@pytest.fixture(scope="session")
def ref_mala():
    lib = ref_lib("libpathref_mala.so")
    if lib is None:
        pytest.skip("oracle/_ref/libpathref_mala.so not built (needs /root/reference; `make ref`)")
    return lib


@pytest.fixture(scope="session")
def ref_hess():
    lib = ref_lib("libpathref_hess.so")
    if lib is None:
        pytest.skip("oracle/_ref/libpathref_hess.so not built (needs /root/reference; `make ref`)")
    return lib


def bsdf_types(c, l, v):
    """BSDF type ids along a serialized path (layout: SURVEY.md App. A.4)."""
    o = 3
    t = []
    if l > 1:
        o += 1 + 56
        for k in range(l - 1):
            t.append(int(v[o + 48]))
            o += 46 + 2 + 10 + (0 if k == l - 2 else 1)
    for k in range(c - 1):
        if k == c - 2:
            if l == 1:
                t.append(int(v[o + 46 + 56]))
            elif l >= 2:
                t.append(int(v[o + 46]))
        else:
            t.append(int(v[o + 48]))
            o += 59
    return t
