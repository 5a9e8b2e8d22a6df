// chain_kernels.cuh -- the persistent-chain kernels (one CUDA thread = one Markov chain) and their
// host launchers.  Instantiated once per MAXD in chain_inst_<MAXD>.cu so the three variants build
// in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "../core/chain.h"

namespace lmc_cuda {
using namespace lmc;

struct DevFilm {
    float *p;
    __device__ __forceinline__ void add(int pix, int c, float v) { atomicAdd(p + 3 * pix + c, v); }
};

// HBM layout of the chain state: word-interleaved ("SoA of 32-bit words"): word w of chain i
// lives at states[w * n + i], so a warp loading / storing its 32 chains touches 32 consecutive
// words per instruction.  Inside the kernel the state is thread-private (local memory, which the
// hardware interleaves per lane the same way).
template <int MAXD>
struct StateWords { static const int NW = (int)(sizeof(ChainState<MAXD>) / 4); };

template <int MAXD>
__device__ __forceinline__ void state_load(const uint32_t *g, int n, int i, ChainState<MAXD> &cs) {
    uint32_t *w = reinterpret_cast<uint32_t *>(&cs);
#pragma unroll 8
    for (int k = 0; k < StateWords<MAXD>::NW; k++) w[k] = g[(size_t)k * n + i];
}
template <int MAXD>
__device__ __forceinline__ void state_store(uint32_t *g, int n, int i, const ChainState<MAXD> &cs) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&cs);
#pragma unroll 8
    for (int k = 0; k < StateWords<MAXD>::NW; k++) g[(size_t)k * n + i] = w[k];
}

template <int MAXD>
__global__ void k_chain_init(uint32_t *states, int n, int chainBase, const float *initLs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ChainState<MAXD> cs;
    memset(&cs, 0, sizeof(cs));
    chain_state_init(cs, initLs ? initLs[chainBase + i] : 0.0f);
    state_store<MAXD>(states, n, i, cs);
}

#ifndef LMC_CHAIN_BLOCK
#define LMC_CHAIN_BLOCK 128
#endif
template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_chain_run(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                uint32_t *states, int n, long long numSteps, float *film,
                                                                unsigned char *trace, float *aTrace) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t tab[64];
    DevFilm df; df.p = film;
    ChainState<MAXD> cs;
    state_load<MAXD>(states, n, i, cs);
    chain_run(sc, rp, chainBase + i, cs, numSteps, tab, 1, df,
              trace ? trace + (size_t)i * numSteps : nullptr, aTrace ? aTrace + (size_t)i * numSteps : nullptr, 1);
    state_store<MAXD>(states, n, i, cs);
}

template <int MAXD>
__global__ void k_chain_stats(const uint32_t *states, int n, unsigned long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[10];
    for (int k = 0; k < 10; k++) v[k] = 0ULL;
    if (i < n) {
        const int o = (int)(offsetof(ChainState<MAXD>, nAccept) / 4);
        for (int k = 0; k < 4; k++) { v[4 + k] = states[(size_t)(o + k) * n + i]; v[k] = states[(size_t)(o + 4 + k) * n + i]; }
        v[8] = states[(size_t)(o + 8) * n + i]; v[9] = states[(size_t)(o + 9) * n + i];
    }
    for (int k = 0; k < 10; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, x);
    }
}


// launchers (defined by LMC_INSTANTIATE_CHAIN in chain_inst_*.cu)
#define LMC_DECLARE_CHAIN(MAXD) \
    size_t chain_state_bytes_##MAXD(); \
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, uint32_t *states, int n, int chainBase, const float *initLs); \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, uint32_t *states, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace); \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const uint32_t *states, int n, unsigned long long *out);
LMC_DECLARE_CHAIN(4)
LMC_DECLARE_CHAIN(8)
LMC_DECLARE_CHAIN(12)

#define LMC_INSTANTIATE_CHAIN(MAXD) \
    size_t chain_state_bytes_##MAXD() { return sizeof(ChainState<MAXD>); } \
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, uint32_t *states, int n, int chainBase, const float *initLs) { \
        k_chain_init<MAXD><<<(n + 127) / 128, 128, 0, st>>>(states, n, chainBase, initLs); \
        return cudaGetLastError(); \
    } \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, uint32_t *states, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace) { \
        k_chain_run<MAXD><<<(n + LMC_CHAIN_BLOCK - 1) / LMC_CHAIN_BLOCK, LMC_CHAIN_BLOCK, 0, st>>>(sc, rp, chainBase, states, n, numSteps, \
                                                                                                film, trace, aTrace); \
        return cudaGetLastError(); \
    } \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const uint32_t *states, int n, unsigned long long *out) { \
        k_chain_stats<MAXD><<<(n + 127) / 128, 128, 0, st>>>(states, n, out); \
        return cudaGetLastError(); \
    }

}  // namespace lmc_cuda
