"""langevin-mcmc_b200 -- host-side Python mirror of the reference's render-loop interface.

The product is `liblmc_b200.so` (hand-written sm_100a CUDA behind the C ABI declared in
include/lmc/lmc_abi.h); this module is a thin ctypes binding that keeps the reference's names
for the path it replaces:

    ParseScene(filename)            src/parsescene.h:8        -> Scene
    Scene.options[...]              src/dptoptions.h:7-34
    MLTInit(scene, ...)             src/mlt.h:41-154          -> (normalization, initLsScore)
    MLT(scene, ...)                 src/mlt.h:156, src/mlt.cpp:20-215 (direct pre-pass + chains + merge [+ WriteImage])
    ChainContext.direct_lighting()  src/direct.cpp:4-54       -> direct sample buffer
    MergeBuffer(...) / WriteImage   src/image.h:79-105, src/image.cpp:29-60
    ChainContext.comm_init / allreduce_film     the film all-reduce of a multi-GPU job (NCCL behind the C ABI)
    MutationType                    src/mutation.h:11

There is NO CPU fallback: importing works without a GPU (so the loader / ABI can be tested),
but every compute call raises LmcError when the CUDA library or a device is missing.
"""
from .api import (LmcError, MutationType, Scene, ParseScene, MLTInit, ChainContext, MLT, load_library,
                  lib_path, decode_trace, MergeBuffer, WriteImage, comm_unique_id, mlt_init_finish)

__all__ = ["LmcError", "MutationType", "Scene", "ParseScene", "MLTInit", "ChainContext", "MLT", "load_library",
           "lib_path", "decode_trace", "MergeBuffer", "WriteImage", "comm_unique_id", "mlt_init_finish"]
