"""ctypes binding of include/lmc/lmc_abi.h (see package docstring)."""
import ctypes
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LmcError(RuntimeError):
    pass


class MutationType(enum.IntEnum):  # src/mutation.h:11
    Large = 0
    Small = 1
    H2MCSmall = 2
    MALASmall = 3


class _SceneInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("width", "height", "num_triangles", "num_bvh_nodes", "num_lights",
                                              "num_shapes", "num_textures", "spp", "direct_spp", "num_init_samples",
                                              "report_interval_spp")]


class _Stats(ctypes.Structure):
    _fields_ = [("proposed", ctypes.c_uint64 * 4), ("accepted", ctypes.c_uint64 * 4),
                ("gradient_evals", ctypes.c_uint64), ("gradient_nonfinite", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint64), ("last_kernel_ms", ctypes.c_double),
                ("outlier_resets", ctypes.c_uint64), ("cache_queries", ctypes.c_uint64), ("cache_hits", ctypes.c_uint64),
                ("cache_count", ctypes.c_uint32 * 5), ("reserved_u32", ctypes.c_uint32)]


class _RunDesc(ctypes.Structure):
    _fields_ = [("num_chains", ctypes.c_int32), ("chain_base", ctypes.c_int32), ("total_chains", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("samples_per_chain", ctypes.c_int64),
                ("normalization", ctypes.c_float), ("reserved_f", ctypes.c_float)]


def lib_path():
    return os.environ.get("LMC_B200_LIB", os.path.join(_HERE, "liblmc_b200.so"))


def load_library():
    """Load liblmc_b200.so; fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise LmcError("CUDA extension %s is missing: run `make lib` (or __graft_entry__.build()); "
                       "there is no CPU fallback" % p)
    L = ctypes.CDLL(p)
    vp, i32, i64, f32, dbl, cp = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double,
                                  ctypes.c_char_p)
    sig = {
        "lmc_last_error": (cp, []), "lmc_version": (cp, []),
        "lmc_scene_load": (i32, [cp, ctypes.POINTER(vp)]), "lmc_scene_save_pack": (i32, [vp, cp]),
        "lmc_scene_free": (None, [vp]), "lmc_scene_get_info": (i32, [vp, ctypes.POINTER(_SceneInfo)]),
        "lmc_scene_set_option": (i32, [vp, cp, dbl]), "lmc_scene_get_option": (i32, [vp, cp, ctypes.POINTER(dbl)]),
        "lmc_scene_serialized": (i32, [vp, vp]),
        "lmc_mlt_init": (i32, [vp, i64, i32, i32, ctypes.POINTER(f32), vp]),
        "lmc_mlt_init_device": (i32, [vp, i64, i32, i32, ctypes.POINTER(f32), vp]),
        "lmc_direct_lighting": (i32, [vp, i32, vp]),
        "lmc_create": (i32, [vp, i32, ctypes.POINTER(vp)]), "lmc_destroy": (None, [vp]),
        "lmc_set_stream": (i32, [vp, vp]),
        "lmc_chains_begin": (i32, [vp, ctypes.POINTER(_RunDesc), vp]),
        "lmc_run_chains": (i32, [vp, i64, vp, vp]), "lmc_synchronize": (i32, [vp]),
        "lmc_get_stats": (i32, [vp, ctypes.POINTER(_Stats)]),
        "lmc_film_clear": (i32, [vp]), "lmc_film_read": (i32, [vp, vp]),
        "lmc_film_device_ptr": (i32, [vp, ctypes.POINTER(vp)]), "lmc_film_bind": (i32, [vp, vp]),
        "lmc_eval_batch": (i32, [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
        "lmc_vert_param_size": (i32, [i32, i32]),
        "lmc_bvh_probe": (i32, [vp, i32, vp, f32, f32, i32, vp, vp, vp]),
        "lmc_create_multi": (i32, [vp, vp, i32, vp]), "lmc_comm_unique_id": (i32, [vp]),
        "lmc_comm_init_rank": (i32, [vp, i32, i32, vp]), "lmc_allreduce_film": (i32, [vp, i32]),
        "lmc_merge_buffer": (i32, [vp, f32, vp, f32, i64, vp]), "lmc_write_image": (i32, [cp, i32, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise LmcError("lmc error %d: %s" % (rc, load_library().lmc_last_error().decode()))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class _Options:
    """Scene.options: the <dpt> block (src/dptoptions.h), addressed by the reference's xml names."""

    def __init__(self, scene):
        self._s = scene

    def __getitem__(self, name):
        v = ctypes.c_double()
        _check(load_library().lmc_scene_get_option(self._s._h, name.encode(), ctypes.byref(v)))
        return v.value

    def __setitem__(self, name, value):
        _check(load_library().lmc_scene_set_option(self._s._h, name.encode(), float(value)))

    def update(self, d):
        for k, v in d.items():
            self[k] = v


class Scene:
    def __init__(self, handle):
        self._h = handle
        self.options = _Options(self)
        info = _SceneInfo()
        _check(load_library().lmc_scene_get_info(self._h, ctypes.byref(info)))
        self.info = {n: getattr(info, n) for n, _ in _SceneInfo._fields_}
        self.width, self.height = info.width, info.height

    def save_pack(self, path):
        _check(load_library().lmc_scene_save_pack(self._h, path.encode()))

    def serialized(self):
        out = np.zeros(38, np.float32)
        _check(load_library().lmc_scene_serialized(self._h, _ptr(out)))
        return out

    def __del__(self):
        try:
            if self._h:
                load_library().lmc_scene_free(self._h)
                self._h = None
        except Exception:
            pass


def ParseScene(filename):
    """src/parsescene.h:8 -- accepts the reference's scene xml (or a pre-flattened .pack)."""
    h = ctypes.c_void_p()
    _check(load_library().lmc_scene_load(os.fspath(filename).encode(), ctypes.byref(h)))
    return Scene(h)


def MLTInit(scene, numInitSamples=None, numChains=None, logicalThreads=32):
    """src/mlt.h:41-154 -> (normalization, initLsScore[numChains])."""
    if numInitSamples is None:
        numInitSamples = scene.info["num_init_samples"]
    if numChains is None:
        numChains = int(scene.options["numchains"])
    norm = ctypes.c_float()
    init_ls = np.zeros(numChains, np.float32)
    _check(load_library().lmc_mlt_init(scene._h, int(numInitSamples), int(numChains), int(logicalThreads),
                                       ctypes.byref(norm), _ptr(init_ls)))
    return norm.value, init_ls


def mlt_init_finish(scores, numInitSamples, numChains):
    """The sequential tail of MLTInit (src/mlt.h:107-153) over the concatenated lsScores of all init contributions
    (ChainContext.mlt_init_part of every rank, in thread order) -> (normalization, initLsScore[numChains])."""
    scores = np.ascontiguousarray(scores, np.float32)
    norm = ctypes.c_float()
    init_ls = np.zeros(int(numChains), np.float32)
    _check(load_library().lmc_mlt_init_finish(_ptr(scores), ctypes.c_int64(scores.size), ctypes.c_int64(int(numInitSamples)),
                                              int(numChains), ctypes.byref(norm), _ptr(init_ls)))
    return norm.value, init_ls


def MergeBuffer(buffer1, b1Weight, buffer2, b2Weight):
    """src/image.h:79-98 followed by BufferToFilm (:100-105): film = b1Weight * buffer1 + b2Weight * buffer2
    (lmc_merge_buffer)."""
    b1 = np.ascontiguousarray(buffer1, np.float32)
    b2 = np.ascontiguousarray(buffer2, np.float32)
    out = np.empty_like(b1)
    _check(load_library().lmc_merge_buffer(_ptr(b1), float(b1Weight), _ptr(b2), float(b2Weight), b1.size, _ptr(out)))
    return out


def WriteImage(filename, film):
    """src/image.cpp:29-60: H x W x 3 float film -> OpenEXR (".exr", three uncompressed float channels) or ".pfm"."""
    film = np.ascontiguousarray(film, np.float32)
    _check(load_library().lmc_write_image(filename.encode(), film.shape[1], film.shape[0], _ptr(film)))


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, every rank passes it to ChainContext.comm_init)."""
    buf = np.zeros(128, np.uint8)
    _check(load_library().lmc_comm_unique_id(_ptr(buf)))
    return buf


def decode_trace(trace):
    """trace byte -> (mutationType, accepted, a>0)."""
    t = np.asarray(trace)
    return t & 3, (t >> 2) & 1, (t >> 3) & 1


class ChainContext:
    """One GPU's chains: lmc_create .. lmc_destroy (the body of MLT()'s ParallelFor, src/mlt.cpp:60-196)."""

    def __init__(self, scene, device=0, stream=None):
        self.scene = scene
        self._c = ctypes.c_void_p()
        _check(load_library().lmc_create(scene._h, int(device), ctypes.byref(self._c)))
        if stream is not None:
            _check(load_library().lmc_set_stream(self._c, ctypes.c_void_p(int(stream))))
        self.num_chains = 0

    def mlt_init(self, numInitSamples, numChains, logicalThreads=65536):
        """MLTInit with the init paths generated on this GPU (lmc_mlt_init_device; src/mlt.h:41-154) ->
        (normalization, initLsScore[numChains]).  Bit-identical to MLTInit(scene, ..., logicalThreads)."""
        norm = ctypes.c_float()
        init_ls = np.zeros(int(numChains), np.float32)
        _check(load_library().lmc_mlt_init_device(self._c, int(numInitSamples), int(numChains), int(logicalThreads),
                                                  ctypes.byref(norm), _ptr(init_ls)))
        return norm.value, init_ls

    def mlt_init_part(self, numInitSamples, logicalThreads, threadBegin, threadEnd):
        """lsScores of the init paths of logical threads [threadBegin, threadEnd), generated on this GPU
        (lmc_mlt_init_device_part): one shard of MLTInit per rank, see mlt_init_finish."""
        n = ctypes.c_int64()
        L = load_library()
        _check(L.lmc_mlt_init_device_part(self._c, ctypes.c_int64(int(numInitSamples)), int(logicalThreads), int(threadBegin),
                                          int(threadEnd), None, ctypes.c_int64(0), ctypes.byref(n)))
        out = np.zeros(max(1, n.value), np.float32)
        _check(L.lmc_mlt_init_device_part(self._c, ctypes.c_int64(int(numInitSamples)), int(logicalThreads), int(threadBegin),
                                          int(threadEnd), _ptr(out), ctypes.c_int64(out.size), ctypes.byref(n)))
        return out[:n.value]

    def direct_lighting(self, direct_spp=None):
        """DirectLighting(scene, buffer), src/direct.cpp:4-54 -> unweighted H x W x 3 sample buffer."""
        if direct_spp is None:
            direct_spp = self.scene.info["direct_spp"]
        out = np.zeros((self.scene.height, self.scene.width, 3), np.float32)
        _check(load_library().lmc_direct_lighting(self._c, int(direct_spp), _ptr(out)))
        return out

    def comm_init(self, nranks, rank, unique_id):
        """Join the film communicator of a one-process-per-GPU job (lmc_comm_init_rank)."""
        uid = np.ascontiguousarray(unique_id, np.uint8)
        _check(load_library().lmc_comm_init_rank(self._c, int(nranks), int(rank), _ptr(uid)))

    def allreduce_film(self):
        """Sum of the films of the job, in place, asynchronous on this ctx's stream (lmc_allreduce_film)."""
        arr = (ctypes.c_void_p * 1)(self._c)
        _check(load_library().lmc_allreduce_film(arr, 1))

    def begin(self, num_chains, normalization, init_ls_score=None, chain_base=0, total_chains=None, *,
              samples_per_chain):
        """samples_per_chain = numSamplesThisChain of the whole run (src/mlt.cpp:64-65): the large-step schedule
        (LS_RATIO, src/mlt.cpp:96) depends on it, so it has no default."""
        d = _RunDesc()
        d.num_chains = int(num_chains)
        d.chain_base = int(chain_base)
        d.total_chains = int(total_chains if total_chains is not None else num_chains)
        d.samples_per_chain = int(samples_per_chain)
        d.normalization = float(normalization)
        if init_ls_score is not None:
            init_ls_score = np.ascontiguousarray(init_ls_score, np.float32)
            if init_ls_score.size != d.total_chains:
                raise LmcError("init_ls_score must have total_chains entries")
        _check(load_library().lmc_chains_begin(self._c, ctypes.byref(d), _ptr(init_ls_score)))
        self.num_chains = d.num_chains

    def run(self, num_mutations, trace=False, a_trace=False):
        t = np.zeros((self.num_chains, num_mutations), np.uint8) if trace else None
        a = np.zeros((self.num_chains, num_mutations), np.float32) if a_trace else None
        _check(load_library().lmc_run_chains(self._c, int(num_mutations), _ptr(t), _ptr(a)))
        return t, a

    def synchronize(self):
        _check(load_library().lmc_synchronize(self._c))

    def stats(self):
        s = _Stats()
        _check(load_library().lmc_get_stats(self._c, ctypes.byref(s)))
        return {"proposed": list(s.proposed), "accepted": list(s.accepted), "gradient_evals": s.gradient_evals,
                "gradient_nonfinite": s.gradient_nonfinite, "kernel_launches": s.kernel_launches,
                "last_kernel_ms": s.last_kernel_ms, "outlier_resets": s.outlier_resets,
                "cache_queries": s.cache_queries, "cache_hits": s.cache_hits, "cache_count": list(s.cache_count)}

    def film(self):
        out = np.zeros((self.scene.height, self.scene.width, 3), np.float32)
        _check(load_library().lmc_film_read(self._c, _ptr(out)))
        return out

    def film_clear(self):
        _check(load_library().lmc_film_clear(self._c))

    def film_device_ptr(self):
        p = ctypes.c_void_p()
        _check(load_library().lmc_film_device_ptr(self._c, ctypes.byref(p)))
        return p.value

    def film_bind(self, device_ptr):
        _check(load_library().lmc_film_bind(self._c, ctypes.c_void_p(int(device_ptr))))

    def eval_batch(self, cam_depth, light_depth, lens, primary, vert_params, want_grad=True, want_hess=False):
        lens = np.ascontiguousarray(lens, np.float32)
        primary = np.ascontiguousarray(primary, np.float32)
        vert_params = np.ascontiguousarray(vert_params, np.float32)
        n = lens.shape[0]
        dim = primary.shape[1] - 1
        log_lum = np.zeros(n, np.float32)
        grad = np.zeros((n, dim), np.float32) if (want_grad or want_hess) else None
        hess = np.zeros((n, dim, dim), np.float32) if want_hess else None
        _check(load_library().lmc_eval_batch(self._c, int(cam_depth), int(light_depth), n, _ptr(lens), _ptr(primary),
                                             _ptr(vert_params), int(vert_params.shape[1]), _ptr(log_lum), _ptr(grad), _ptr(hess)))
        if want_hess:
            return log_lum, grad, hess
        return log_lum, grad

    def bvh_probe(self, rays, tmin, tmax, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]
        tid = np.zeros(n, np.int32)
        gp = np.zeros((n, 2), np.int32)
        tuv = np.zeros((n, 3), np.float32)
        _check(load_library().lmc_bvh_probe(self._c, n, _ptr(rays), float(tmin), float(tmax), 1 if any_hit else 0,
                                            _ptr(tid), _ptr(gp), _ptr(tuv)))
        return tid, gp, tuv

    def close(self):
        if self._c:
            load_library().lmc_destroy(self._c)
            self._c = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def MLT(scene, numChains=None, mutationsPerChain=None, device=0, logicalThreads=32, numInitSamples=None, directSpp=None,
        outputName=None):
    """MLT() (src/mlt.cpp:20-215) on one GPU: DirectLighting pre-pass, MLTInit on the host, the chain loop on the
    GPU, MergeBuffer(direct / directSpp, indirect / spp) -> film, optional WriteImage.  Returns (film H x W x 3,
    stats).  Known deviations from the reference's MLT(): every chain runs mutationsPerChain iterations (the
    reference adds one to chainId < numSamplesPerChain % numChains, src/mlt.cpp:40,64-65); the global cache
    (src/mutation_mala.h:131-161) is an option here (`globalcache`, default 0 = every eligible MALA step evaluates its gradient;
    1 = cached moments once a dimension holds 3000 entries, filled in a defined order instead of the reference's thread order);
    a gradient that is not evaluated because ssScore <= 1e-10 is zero instead of the stale vector the reference
    reuses.  The C++ form with multi-GPU sharding and progressive dumps is include/lmc/mlt.h."""
    if numChains is None:
        numChains = int(scene.options["numchains"])
    if mutationsPerChain is None:
        mutationsPerChain = scene.info["spp"] * scene.width * scene.height // numChains
    if directSpp is None:
        directSpp = scene.info["direct_spp"]
    norm, init_ls = MLTInit(scene, numInitSamples, numChains, logicalThreads)
    ctx = ChainContext(scene, device)
    direct = ctx.direct_lighting(directSpp) if directSpp > 0 else np.zeros((scene.height, scene.width, 3), np.float32)
    ctx.begin(numChains, norm, init_ls, samples_per_chain=mutationsPerChain)
    ctx.run(mutationsPerChain)
    indirect = ctx.film()
    stats = ctx.stats()
    ctx.close()
    spp = numChains * mutationsPerChain / float(scene.width * scene.height)
    film = MergeBuffer(direct, 1.0 / directSpp if directSpp > 0 else 0.0, indirect, 1.0 / spp if spp > 0 else 0.0)
    if outputName:
        WriteImage(outputName, film)
    return film, stats
