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

// ---- wavefront execution of one chain-loop iteration -----------------------------------------
// The iteration of src/mlt.cpp:91-170 is cut into the phases of core/mutation.h; every phase
// is its own kernel so that (i) each kernel's instruction footprint is a fraction of the whole
// loop body, (ii) divergent work (large steps, gradient evaluations) runs on COMPACTED chain
// lists with full warps.  Lists are filled with warp-aggregated atomics by the preceding phase.
struct WaveLists {
    int *large, *small_, *curGrad, *propGrad;   // chain slots (local ids), capacity n each
    int *counts;                                // [0] large [1] small [2] curGrad [3] propGrad
};

__device__ __forceinline__ void list_append(int *list, int *counter, bool pred, int value) {
    const unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    list[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

template <int MAXD>
__device__ __forceinline__ void rng_open(Rng &rng, uint32_t *tab, const Scene &sc, int globalChainId, ChainState<MAXD> &cs) {
    rng.tab = tab; rng.stride = 1;
    const uint64_t seed = (uint64_t)(long long)(globalChainId + sc.opt.seedOffset);
    if (!cs.seeded) { rng_seed(rng, seed); cs.seeded = 1u; }
    else rng_restore(rng, seed, cs.rngState, cs.rngEpoch);
}
template <int MAXD>
__device__ __forceinline__ void rng_close(const Rng &rng, ChainState<MAXD> &cs) { cs.rngState = rng.state; cs.rngEpoch = rng.epoch; }

template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_wave_begin(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                 uint32_t *states, int n, WaveLists wl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    int kind = -1, needCur = 0;
    if (active) {
        uint32_t tab[64];
        ChainState<MAXD> cs;
        state_load<MAXD>(states, n, i, cs);
        Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
        phase_begin(sc, rp, cs.sampleIdx, cs.st[cs.curIdx], cs.ch, rng, cs.ss);
        rng_close(rng, cs);
        state_store<MAXD>(states, n, i, cs);
        kind = cs.ss.kind; needCur = cs.ss.needCurGrad;
    }
    list_append(wl.large, wl.counts + 0, active && kind == STEP_LARGE, i);
    list_append(wl.small_, wl.counts + 1, active && kind != STEP_LARGE, i);
    list_append(wl.curGrad, wl.counts + 2, active && needCur, i);
}

// gradient of the current state (which = 0) or of the proposal (which = 1) for a compacted list
template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_wave_grad(const __grid_constant__ Scene sc, uint32_t *states, int n,
                                                                const int *list, const int *count, int which) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    const int i = list[t];
    ChainState<MAXD> cs;
    state_load<MAXD>(states, n, i, cs);
    phase_gradient(sc, cs.st[cs.curIdx ^ which], cs.ss, cs.gradStats);
    state_store<MAXD>(states, n, i, cs);
}

template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_wave_propose(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                   uint32_t *states, int n, const int *list, const int *count,
                                                                   WaveLists wl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < *count;
    int i = -1, needProp = 0;
    if (active) {
        i = list[t];
        uint32_t tab[64];
        ChainState<MAXD> cs;
        state_load<MAXD>(states, n, i, cs);
        Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
        phase_propose(sc, rp, cs.st[cs.curIdx], cs.st[cs.curIdx ^ 1], cs.ch, rng, cs.ss);
        rng_close(rng, cs);
        state_store<MAXD>(states, n, i, cs);
        needProp = cs.ss.needPropGrad;
    }
    list_append(wl.propGrad, wl.counts + 3, active && needProp, i);
}

template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_wave_finish(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                  uint32_t *states, int n, float *film, unsigned char *trace,
                                                                  float *aTrace, long long numSteps, long long stepInLaunch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t tab[64];
    DevFilm df; df.p = film;
    ChainState<MAXD> cs;
    state_load<MAXD>(states, n, i, cs);
    Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
    const StepInfo info = phase_finish(sc, rp, chainBase + i, cs.sampleIdx, cs.st, cs.curIdx, cs.ch, rng, df, cs.ss);
    rng_close(rng, cs);
    cs.nPropose[info.mutationType] += 1u;
    cs.nAccept[info.mutationType] += (unsigned int)info.accepted;
    cs.sampleIdx += 1;
    state_store<MAXD>(states, n, i, cs);
    if (trace) trace[(size_t)i * numSteps + stepInLaunch] = (unsigned char)(info.mutationType | (info.accepted << 2) | ((info.a > 0.0f) ? 8 : 0));
    if (aTrace) aTrace[(size_t)i * numSteps + stepInLaunch] = info.a;
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
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, unsigned long long *launches); \
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
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, unsigned long long *launches) { \
        const int B = LMC_CHAIN_BLOCK, G = (n + B - 1) / B; \
        for (long long k = 0; k < numSteps; k++) { \
            cudaError_t e = cudaMemsetAsync(wl.counts, 0, 4 * sizeof(int), st); \
            if (e != cudaSuccess) return e; \
            k_wave_begin<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl); \
            k_wave_grad<MAXD><<<G, B, 0, st>>>(sc, states, n, wl.curGrad, wl.counts + 2, 0); \
            k_wave_propose<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.small_, wl.counts + 1, wl); \
            k_wave_propose<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.large, wl.counts + 0, wl); \
            k_wave_grad<MAXD><<<G, B, 0, st>>>(sc, states, n, wl.propGrad, wl.counts + 3, 1); \
            k_wave_finish<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, film, trace, aTrace, numSteps, k); \
            *launches += 6; \
            e = cudaGetLastError(); \
            if (e != cudaSuccess) return e; \
        } \
        return cudaSuccess; \
    } \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const uint32_t *states, int n, unsigned long long *out) { \
        k_chain_stats<MAXD><<<(n + 127) / 128, 128, 0, st>>>(states, n, out); \
        return cudaGetLastError(); \
    }

}  // namespace lmc_cuda
