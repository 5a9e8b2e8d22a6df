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

// HBM layout of the chain state: one 16-byte aligned record per chain (AoS).  A thread moves
// (parts of) its record with 16-byte loads/stores; consecutive 16-byte pieces of one record
// share 128-byte lines, so the traffic is sector-exact even when a warp's chains are a
// PERMUTED set (the sorted work lists below), which a lane-interleaved layout would not survive.
template <int MAXD>
struct alignas(16) ChainRec { ChainState<MAXD> cs; };

__device__ __forceinline__ void copy16(void *dst, const void *src, int bytes) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    const int n16 = bytes >> 4;
#pragma unroll 4
    for (int k = 0; k < n16; k++) d[k] = s[k];
    // tail (records are padded to 16 B, ranges may not be)
    unsigned char *db = reinterpret_cast<unsigned char *>(dst) + (n16 << 4);
    const unsigned char *sb = reinterpret_cast<const unsigned char *>(src) + (n16 << 4);
    for (int k = 0; k < (bytes & 15); k++) db[k] = sb[k];
}
template <int MAXD>
__device__ __forceinline__ void state_load(const ChainRec<MAXD> *g, int i, ChainRec<MAXD> &r) { copy16(&r, g + i, (int)sizeof(ChainRec<MAXD>)); }
template <int MAXD>
__device__ __forceinline__ void state_store(ChainRec<MAXD> *g, int i, const ChainRec<MAXD> &r) { copy16(g + i, &r, (int)sizeof(ChainRec<MAXD>)); }

template <int MAXD>
__global__ void k_chain_init(ChainRec<MAXD> *states, int n, int chainBase, const float *initLs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ChainRec<MAXD> r;
    memset(&r, 0, sizeof(r));
    chain_state_init(r.cs, initLs ? initLs[chainBase + i] : 0.0f);
    state_store<MAXD>(states, i, r);
}

#ifndef LMC_CHAIN_BLOCK
#define LMC_CHAIN_BLOCK 128
#endif

// ---- wavefront execution of one chain-loop iteration -----------------------------------------
// The iteration of src/mlt.cpp:91-170 is cut into the phases of core/mutation.h; every phase
// is its own kernel so that (i) each kernel's instruction footprint is a fraction of the whole
// loop body, (ii) divergent work runs on COMPACTED chain lists with full warps, and (iii) the
// lists of the expensive phases are SORTED by path class (camDepth, lightDepth, step kind) with
// an on-device counting sort, so the warps of a block walk the same control flow.
#define LMC_NKEYS 256
struct SortList {
    int *keys;      // n: class key of each chain for this list, or -1 = not in the list
    int *hist;      // LMC_NKEYS (zero between uses)
    int *offsets;   // LMC_NKEYS
    int *cursor;    // LMC_NKEYS
    int *list;      // n
    int *count;     // 1
};
struct WaveLists {
    SortList small_, curGrad, propGrad;
    int *large, *largeCount;     // unsorted
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
__device__ __forceinline__ void sort_key_set(const SortList &sl, int i, int key) {
    sl.keys[i] = key;
    if (key >= 0) atomicAdd(sl.hist + key, 1);
}
__device__ __forceinline__ int class_key(int camDepth, int lgtDepth, int kindBit) {
    int k = ((camDepth & 15) * 9 + (lgtDepth < 8 ? lgtDepth : 8)) * 2 + kindBit;
    return k < LMC_NKEYS ? k : LMC_NKEYS - 1;
}

// 1 block: exclusive scan of up to 3 histograms; clears hist + cursor for the next use
static __global__ void k_sort_scan(SortList a, SortList b, int nb) {
    SortList sl = (blockIdx.x == 0) ? a : b;
    if ((int)blockIdx.x >= nb) return;
    __shared__ int sh[LMC_NKEYS];
    const int t = threadIdx.x;
    sh[t] = sl.hist[t];
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int k = 0; k < LMC_NKEYS; k++) { const int c = sh[k]; sh[k] = acc; acc += c; }
        *sl.count = acc;
    }
    __syncthreads();
    sl.offsets[t] = sh[t];
    sl.hist[t] = 0;
    sl.cursor[t] = 0;
}
static __global__ void k_sort_scatter(int n, SortList a, SortList b, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ka = a.keys[i];
    if (ka >= 0) a.list[a.offsets[ka] + atomicAdd(a.cursor + ka, 1)] = i;
    if (nb > 1) {
        const int kb = b.keys[i];
        if (kb >= 0) b.list[b.offsets[kb] + atomicAdd(b.cursor + kb, 1)] = i;
    }
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
                                                                 ChainRec<MAXD> *states, int n, WaveLists wl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    int kind = -1;
    if (active) {
        uint32_t tab[64];
        ChainState<MAXD> &cs = states[i].cs;     // phases touch a few sectors of the record: work in place
        Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
        phase_begin(sc, rp, cs.sampleIdx, cs.st[cs.curIdx], cs.ch, rng, cs.ss);
        rng_close(rng, cs);
        kind = cs.ss.kind;
        const Path<MAXD> &p = cs.st[cs.curIdx].path;
        sort_key_set(wl.small_, i, kind == STEP_LARGE ? -1 : class_key(p.camDepth, p.lgtDepth, kind == STEP_ISO ? 0 : 1));
        sort_key_set(wl.curGrad, i, cs.ss.needCurGrad ? class_key(p.camDepth, p.lgtDepth, 0) : -1);
    }
    list_append(wl.large, wl.largeCount, active && kind == STEP_LARGE, i);
}

// gradient of the current state (which = 0) or of the proposal (which = 1) for a sorted list
#ifndef LMC_GRAD_MINB
#define LMC_GRAD_MINB 2
#endif
#ifndef LMC_PROP_MINB
#define LMC_PROP_MINB 4
#endif
template <int MAXD, int ORDER>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK, LMC_GRAD_MINB) k_wave_grad(const __grid_constant__ Scene sc, ChainRec<MAXD> *states, int n,
                                                                const int *list, const int *count, int which, H2mcSide *sides) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    const int i = list[t];
    ChainState<MAXD> &cs = states[i].cs;
    phase_gradient<MAXD, ORDER>(sc, cs.st[cs.curIdx ^ which], cs.ss, cs.gradStats, sides ? sides + i : nullptr);
}

template <int MAXD, int ONLY>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK, LMC_PROP_MINB) k_wave_propose(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                   ChainRec<MAXD> *states, int n, const int *list, const int *count,
                                                                   WaveLists wl, H2mcSide *sides) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    const int i = list[t];
    uint32_t tab[64];
    ChainState<MAXD> &cs = states[i].cs;
    Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
    phase_propose<MAXD, ONLY>(sc, rp, cs.st[cs.curIdx], cs.st[cs.curIdx ^ 1], cs.ch, rng, cs.ss, sides ? sides + i : nullptr, cs.curIdx);
    rng_close(rng, cs);
    const MarkovState<MAXD> &prop = cs.st[cs.curIdx ^ 1];
    sort_key_set(wl.propGrad, i, cs.ss.needPropGrad ? class_key(prop.sp.camDepth, prop.sp.lightDepth, 0) : -1);
}

template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_wave_finish(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                  ChainRec<MAXD> *states, int n, float *film, unsigned char *trace,
                                                                  float *aTrace, long long numSteps, long long stepInLaunch, H2mcSide *sides) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t tab[64];
    DevFilm df; df.p = film;
    ChainState<MAXD> &cs = states[i].cs;
    Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
    const StepInfo info = phase_finish(sc, rp, chainBase + i, cs.sampleIdx, cs.st, cs.curIdx, cs.ch, rng, df, cs.ss, sides ? sides + i : nullptr);
    rng_close(rng, cs);
    cs.nPropose[info.mutationType] += 1u;
    cs.nAccept[info.mutationType] += (unsigned int)info.accepted;
    cs.sampleIdx += 1;
    if (trace) trace[(size_t)i * numSteps + stepInLaunch] = (unsigned char)(info.mutationType | (info.accepted << 2) | ((info.a > 0.0f) ? 8 : 0));
    if (aTrace) aTrace[(size_t)i * numSteps + stepInLaunch] = info.a;
}

template <int MAXD>
__global__ void k_chain_stats(const ChainRec<MAXD> *states, int n, unsigned long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[10];
    for (int k = 0; k < 10; k++) v[k] = 0ULL;
    if (i < n) {
        const ChainState<MAXD> &cs = states[i].cs;
        for (int k = 0; k < 4; k++) { v[k] = cs.nPropose[k]; v[4 + k] = cs.nAccept[k]; }
        v[8] = cs.gradStats[0]; v[9] = cs.gradStats[1];
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
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, void *states, int n, int chainBase, const float *initLs); \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, void *states, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, unsigned long long *launches, H2mcSide *sides); \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const void *states, int n, unsigned long long *out);
LMC_DECLARE_CHAIN(4)
LMC_DECLARE_CHAIN(8)
LMC_DECLARE_CHAIN(12)

#define LMC_INSTANTIATE_CHAIN(MAXD) \
    size_t chain_state_bytes_##MAXD() { return sizeof(ChainRec<MAXD>); } \
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, void *states, int n, int chainBase, const float *initLs) { \
        k_chain_init<MAXD><<<(n + 127) / 128, 128, 0, st>>>((ChainRec<MAXD> *)states, n, chainBase, initLs); \
        return cudaGetLastError(); \
    } \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, void *states_, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, unsigned long long *launches, H2mcSide *sides) { \
        ChainRec<MAXD> *states = (ChainRec<MAXD> *)states_; \
        const int B = LMC_CHAIN_BLOCK, G = (n + B - 1) / B; \
        for (long long k = 0; k < numSteps; k++) { \
            cudaError_t e = cudaMemsetAsync(wl.largeCount, 0, sizeof(int), st); \
            if (e != cudaSuccess) return e; \
            k_wave_begin<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl); \
            k_sort_scan<<<2, LMC_NKEYS, 0, st>>>(wl.small_, wl.curGrad, 2); \
            k_sort_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, wl.small_, wl.curGrad, 2); \
            if (sc.opt.h2mc) k_wave_grad<MAXD, 2><<<G, B, 0, st>>>(sc, states, n, wl.curGrad.list, wl.curGrad.count, 0, sides); \
            else k_wave_grad<MAXD, 1><<<G, B, 0, st>>>(sc, states, n, wl.curGrad.list, wl.curGrad.count, 0, sides); \
            k_wave_propose<MAXD, 1><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.small_.list, wl.small_.count, wl, sides); \
            k_wave_propose<MAXD, 0><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.large, wl.largeCount, wl, sides); \
            k_sort_scan<<<1, LMC_NKEYS, 0, st>>>(wl.propGrad, wl.propGrad, 1); \
            k_sort_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, wl.propGrad, wl.propGrad, 1); \
            if (sc.opt.h2mc) k_wave_grad<MAXD, 2><<<G, B, 0, st>>>(sc, states, n, wl.propGrad.list, wl.propGrad.count, 1, sides); \
            else k_wave_grad<MAXD, 1><<<G, B, 0, st>>>(sc, states, n, wl.propGrad.list, wl.propGrad.count, 1, sides); \
            k_wave_finish<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, film, trace, aTrace, numSteps, k, sides); \
            *launches += 10; \
            e = cudaGetLastError(); \
            if (e != cudaSuccess) return e; \
        } \
        return cudaSuccess; \
    } \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const void *states, int n, unsigned long long *out) { \
        k_chain_stats<MAXD><<<(n + 127) / 128, 128, 0, st>>>((const ChainRec<MAXD> *)states, n, out); \
        return cudaGetLastError(); \
    }

}  // namespace lmc_cuda
