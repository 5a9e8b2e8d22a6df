// film_comm.h -- the one collective of the chain phase: sum of the fp32 film over the GPUs of a job.
//
// The reference is a single process whose threads splat into one shared SampleBuffer (src/mlt.cpp:55,
// src/image.h:66-77).  Sharded over GPUs, every device accumulates its own W*H*3 film and one
// ncclAllReduce(sum, fp32) over NVLink / NVSwitch at a checkpoint (end of run, or a progressive dump every
// reportIntervalSpp, src/mlt.cpp:171-193) restores the shared buffer.  Two ways to form the
// communicator, both behind the C ABI:
//   one process, several GPUs   lmc_create_multi   -> ncclCommInitAll + grouped all-reduce
//   one process per GPU         lmc_comm_unique_id / lmc_comm_init_rank  (the id travels by any side channel)
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host framework already loaded, else the system
// one), so liblmc_b200.so has no link-time dependency on it and single-GPU users never touch it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string>

namespace lmc_cuda {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool load() {
        if (lib) return true;
        const char *names[] = {getenv("LMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { error = "NCCL not found (libnccl.so.2); set LMC_NCCL_LIB"; return false; }
#define LMC_NCCL_SYM(field, sym) field = (decltype(field))dlsym(lib, sym); if (!field) { error = std::string("NCCL symbol missing: ") + sym; lib = nullptr; return false; }
        LMC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        LMC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        LMC_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        LMC_NCCL_SYM(AllReduce, "ncclAllReduce")
        LMC_NCCL_SYM(GroupStart, "ncclGroupStart")
        LMC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        LMC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        LMC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LMC_NCCL_SYM
        return true;
    }
};

inline NcclApi &nccl_api() { static NcclApi api; return api; }

}  // namespace lmc_cuda
