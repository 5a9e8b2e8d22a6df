// chain_kernels.cuh -- the kernels of the chain loop (src/mlt.cpp:91-170) and their host launcher.
//
//   bookkeeping phases   k_wave_begin, k_wave_finish<THEN_BEGIN>, k_prop_start, k_prop_post_large   (thread = chain)
//   work lists           k_sort_scan, k_sort_scatter (counting sort by path class / screen tile)
//   gradient / Hessian   k_wave_grad<ORDER> (class-pure blocks in lockstep)
//   proposal, per-vertex wavefront (>= ~4e5 chains):  k_trace / k_shadow (trace_kernels.cuh), k_shade<stage>,
//                        k_shade_tail, k_connect        (thread = pending ray / queue entry)
//   proposal, monolithic (small jobs):                k_wave_propose<ONLY>
//   set-up               k_chain_init, k_mlt_init_paths; k_chain_stats
// launch_chain_run_t() is one lmc_run_chains call.  Instantiated once per MAXD in chain_inst_<MAXD>.cu so
// the three variants build in parallel.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../core/chain.h"
#include "trace_kernels.cuh"

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
#ifndef LMC_FINISH_MINB
#define LMC_FINISH_MINB 5          // resident blocks per SM asked of k_wave_finish: <= 102 registers, 20 warps (measured: 4..6 are
                                   // equivalent, 8 = 64 registers is 3 % slower, and 1 lets ptxas take 168 registers: +0.12 ms)
#endif

// ---- wavefront execution of one chain-loop iteration -----------------------------------------
// The iteration of src/mlt.cpp:91-170 is cut into the phases of core/mutation.h; every phase
// is its own kernel so that (i) each kernel's instruction footprint is a fraction of the whole
// loop body, (ii) divergent work runs on COMPACTED chain lists with full warps, and (iii) the
// lists of the expensive phases are SORTED by path class (camDepth, lightDepth, step kind) with
// an on-device counting sort, so the warps of a block walk the same control flow.
#define LMC_NKEYS 256            // path classes
#ifndef LMC_TILES
#define LMC_TILES 16             // the small-step list is also sorted by screen tile (LMC_TILES x LMC_TILES)
#endif
#define LMC_NKEYS_SMALL (LMC_NKEYS * LMC_TILES * LMC_TILES)
struct SortList {
    int nkeys;      // LMC_NKEYS, or LMC_NKEYS_SMALL for the small-step list
    int *keys;      // n: key of each chain for this list, or -1 = not in the list
    int *hist;      // nkeys (zero between uses)
    int *offsets;   // nkeys
    int *cursor;    // nkeys
    int *list;      // n
    int *count;     // 1
};
struct WaveLists {
    int listLen;                 // capacity of every sorted list: n + room for class-alignment gaps
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
// counter++ for every thread that gets here, as ONE atomic per group of converged lanes: the shadow-ray and connection
// queues take millions of pushes per iteration on a single counter, and an atomic that returns its value costs the
// pushing warp a full L2 round trip per lane (ncu, veach-door k_shade<G_CAM>: 30 % of all stall samples sat behind the
// two ATOM instructions of the per-thread form).
__device__ __forceinline__ int warp_agg_inc(int *counter) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}
__device__ __forceinline__ void sort_key_set(const SortList &sl, int i, int key) {
    sl.keys[i] = key;
    if (key < 0) return;
    if (sl.nkeys <= LMC_NKEYS) {
        // the gradient lists have a few dozen class keys: lanes with the same key add once.  (Not for the small-step list:
        // its 65 536 (class, tile) keys are mostly distinct within a warp and __match_any_sync costs a round per distinct
        // value -- measured +0.12 ms on k_wave_finish.)
        const unsigned peers = __match_any_sync(__activemask(), key);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(sl.hist + key, __popc(peers));
    } else atomicAdd(sl.hist + key, 1);
}
__device__ __forceinline__ int class_key(int camDepth, int lgtDepth, int kindBit) {
    int k = ((camDepth & 15) * 9 + (lgtDepth < 8 ? lgtDepth : 8)) * 2 + kindBit;
    return k < LMC_NKEYS ? k : LMC_NKEYS - 1;
}

// 1 block (1024 threads) per list: exclusive scan of the key histogram; clears hist + cursor for the
// next use.  align > 1 (class-keyed lists only) starts every class at a multiple of `align` (class-pure
// thread blocks for the gradient kernel); the gaps keep the -1 the list was pre-filled with, and *count is
// the padded length.
static __global__ void __launch_bounds__(1024) k_sort_scan(SortList a, SortList b, int nb, int alignA, int alignB) {
    SortList sl = (blockIdx.x == 0) ? a : b;
    const int align = (blockIdx.x == 0) ? alignA : alignB;
    if ((int)blockIdx.x >= nb) return;
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int nkeys = sl.nkeys;
    if (align > 1) {            // few keys (<= 1024), non-associative rounding: serial over shared memory
        if (t < nkeys) part[t] = sl.hist[t];
        __syncthreads();
        if (t == 0) {
            int acc = 0;
            for (int k = 0; k < nkeys; k++) {
                const int c = part[k];
                if (c > 0) acc = (acc + align - 1) / align * align;
                part[k] = acc; acc += c;
            }
            acc = (acc + align - 1) / align * align;
            *sl.count = acc;
        }
        __syncthreads();
        if (t < nkeys) { sl.offsets[t] = part[t]; sl.hist[t] = 0; sl.cursor[t] = 0; }
        return;
    }
    // tiles of 1024 keys: coalesced loads, warp-shuffle scan, running offset carried from tile to tile
    __shared__ int running;
    if (t == 0) running = 0;
    __syncthreads();
    const int lane = t & 31, warp = t >> 5;
    for (int k0 = 0; k0 < nkeys; k0 += 1024) {
        const int k = k0 + t;
        const int c = (k < nkeys) ? sl.hist[k] : 0;
        int v = c;
        for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += u; }
        if (lane == 31) part[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = part[lane];
            for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += u; }
            part[32 + lane] = w;          // inclusive scan of the warp totals
        }
        __syncthreads();
        const int base = running + (warp > 0 ? part[32 + warp - 1] : 0);
        if (k < nkeys) { sl.offsets[k] = base + v - c; sl.hist[k] = 0; sl.cursor[k] = 0; }
        __syncthreads();
        if (t == 0) running += part[32 + 31];
        __syncthreads();
    }
    if (t == 0) *sl.count = running;
}
// one atomic per (warp, key): lanes holding the same key reserve a run of slots together
__device__ __forceinline__ void sort_scatter_one(const SortList &sl, int i, bool inRange) {
    const int key = inRange ? sl.keys[i] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key < 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(sl.cursor + key, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    sl.list[sl.offsets[key] + base + __popc(peers & ((1u << lane) - 1u))] = i;
}
static __global__ void k_sort_scatter(int n, SortList a, SortList b, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    sort_scatter_one(a, i, i < n);
    if (nb > 1) sort_scatter_one(b, i, i < n);
}

template <int MAXD>
__device__ __forceinline__ void rng_open(Rng &rng, uint32_t *tab, const Scene &sc, int globalChainId, ChainState<MAXD> &cs) {
    // lazy table: `tab` (local scratch) is written only if the table must advance (2^-32 per draw)
    rng.tab = tab; rng.stride = 1;
    const uint64_t seed = (uint64_t)(long long)(globalChainId + sc.opt.seedOffset);
    if (!cs.seeded) { rng_seed_lazy(rng, seed); cs.seeded = 1u; }
    else rng_restore_lazy(rng, seed, cs.rngState, cs.rngEpoch);
}
template <int MAXD>
__device__ __forceinline__ void rng_close(const Rng &rng, ChainState<MAXD> &cs) { cs.rngState = rng.state; cs.rngEpoch = rng.epoch; }

// phase_begin + work-list keys for chain i (shared by k_wave_begin and the fused k_wave_finish)
template <int MAXD>
__device__ __forceinline__ int wave_begin_chain(const Scene &sc, const RunParams &rp, ChainState<MAXD> &cs, Rng &rng, int i, const WaveLists &wl) {
    phase_begin(sc, rp, cs.sampleIdx, cs.st[cs.curIdx], cs.ch, rng, cs.ss);
    const int kind = cs.ss.kind;
    const Path<MAXD> &p = cs.st[cs.curIdx].path;
    // small steps: (class, screen tile) -- chains of a warp then perturb paths through the same part of the
    // scene: coherent primary rays, neighbouring hit points, the same materials
    const int tx = dm_clampi((int)(p.screenPos.x * (float)LMC_TILES), 0, LMC_TILES - 1);
    const int ty = dm_clampi((int)(p.screenPos.y * (float)LMC_TILES), 0, LMC_TILES - 1);
    sort_key_set(wl.small_, i, kind == STEP_LARGE ? -1 : class_key(p.camDepth, p.lgtDepth, kind == STEP_ISO ? 0 : 1) * (LMC_TILES * LMC_TILES) + ty * LMC_TILES + tx);
    sort_key_set(wl.curGrad, i, cs.ss.needCurGrad ? class_key(p.camDepth, p.lgtDepth, 0) : -1);
    return kind;
}

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
        kind = wave_begin_chain(sc, rp, cs, rng, i, wl);
        rng_close(rng, cs);
    }
    list_append(wl.large, wl.largeCount, active && kind == STEP_LARGE, i);
}

}  // namespace lmc_cuda
#include "h2mc_kernels.cuh"
namespace lmc_cuda {

// gradient of the current state (which = 0) or of the proposal (which = 1) for a sorted list
// occupancy of the gradient kernel in units of 128 threads per SM.  The reverse-sweep evaluator is bound by the
// latency of its local-memory traffic (checkpoints, spills): 4 (two 256-thread blocks, 128 registers) measured
// 1.69 ms per proposal-gradient launch against 2.28 ms at 2 (255 registers) on 2^20 chains (profiles/r02_*).
#ifndef LMC_GRAD_MINB
#define LMC_GRAD_MINB 4
#endif
#ifndef LMC_PROP_MINB
#define LMC_PROP_MINB 4
#endif
#ifndef LMC_GRAD_BLOCK
#define LMC_GRAD_BLOCK 256
#endif
// The gradient lists are class-aligned to LMC_GRAD_BLOCK (k_sort_scan), so every block evaluates paths
// of ONE (camDepth, lightDepth) class: the evaluator's vertex loops have the same trip counts in all of
// its threads, which is what the barriers of core/pathgrad.h (LMC_VERTEX_SYNC) require.  Padding
// entries (-1) redo the block's first chain into scratch so that they take part in every barrier.
// (the Hessian instantiation, ORDER = 2, carries dual numbers through the sweep and has its own occupancy setting)
#ifndef LMC_HESS_MINB
#define LMC_HESS_MINB 4
#endif
#ifndef LMC_HESS_BLOCK
#define LMC_HESS_BLOCK LMC_GRAD_BLOCK      // must divide LMC_GRAD_BLOCK (the class alignment of the gradient lists)
#endif
template <int MAXD, int ORDER>
__global__ void __launch_bounds__(ORDER == 2 ? LMC_HESS_BLOCK : LMC_GRAD_BLOCK,
                                  ORDER == 2 ? (LMC_HESS_MINB * 128) / LMC_HESS_BLOCK : (LMC_GRAD_MINB * 128) / LMC_GRAD_BLOCK) k_wave_grad(const __grid_constant__ Scene sc, ChainRec<MAXD> *states, int n,
                                                                const int *list, const int *count, int which, H2mcSide *sides,
                                                                H2mcSide *padSide) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;                    // *count is a multiple of the block size: whole blocks leave
    int i = list[t];
    const bool pad = i < 0;
    if (pad) {
        i = list[blockIdx.x * blockDim.x];      // class segments start on alignment boundaries and are filled from the front
        if (i < 0) return;                      // (a block smaller than the alignment can be all padding: uniform exit)
    }
    ChainState<MAXD> &cs = states[i].cs;
    if (pad) {
        StepScratch<MAXD> scratch; scratch.kind = cs.ss.kind;
        phase_gradient<MAXD, ORDER>(sc, cs.st[cs.curIdx ^ which], scratch, nullptr, padSide);
        return;
    }
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
    sort_key_set(wl.propGrad, i, cs.ss.needPropGrad ? class_key(prop.path.camDepth, prop.path.lgtDepth, 0) : -1);
}

// Phase 4 of iteration k and, when THEN_BEGIN, phase 0 of iteration k + 1 in the same pass over the records
template <int MAXD, int THEN_BEGIN>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK, LMC_FINISH_MINB) k_wave_finish(const __grid_constant__ Scene sc, RunParams rp, int chainBase,
                                                                  ChainRec<MAXD> *states, int n, float *film, unsigned char *trace,
                                                                  float *aTrace, long long numSteps, long long stepInLaunch, H2mcSide *sides,
                                                                  WaveLists wl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    int kind = -1;
    if (active) {
        uint32_t tab[64];
        DevFilm df; df.p = film;
        ChainState<MAXD> &cs = states[i].cs;
        Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
        const StepInfo info = phase_finish(sc, rp, chainBase + i, cs.sampleIdx, cs.st, cs.curIdx, cs.ch, rng, df, cs.ss, sides ? sides + i : nullptr);
        cs.nPropose[info.mutationType] += 1u;
        cs.nAccept[info.mutationType] += (unsigned int)info.accepted;
        cs.sampleIdx += 1;
        if (trace) trace[(size_t)i * numSteps + stepInLaunch] = (unsigned char)(info.mutationType | (info.accepted << 2) | ((info.a > 0.0f) ? 8 : 0));
        if (aTrace) aTrace[(size_t)i * numSteps + stepInLaunch] = info.a;
        if (THEN_BEGIN) kind = wave_begin_chain(sc, rp, cs, rng, i, wl);
        rng_close(rng, cs);
    }
    if (THEN_BEGIN) list_append(wl.large, wl.largeCount, active && kind == STEP_LARGE, i);
}

// ---- per-vertex wavefront of the proposal phase ------------------------------------------------
// The proposal (PerturbPathBidir / GeneratePathBidir) advances one path vertex per WAVE:
//   k_trace      closest hit for every pending ray of the wave
//   k_shade<S>   the reference's statements between two ray queries (core/stages.h) for the chains
//                whose pending ray belongs to stage S; pushes the next ray into the other queue
//                set and the shadow rays of candidate contributions into the shadow queue
// and after the last wave
//   k_shadow     any-hit for all queued connection segments -> candidate flags
//   k_prop_post_large  contribution choice / acceptance bookkeeping of the large steps (small steps: post_small_now)
//
// A queue entry carries the whole between-stage state of its proposal (TraceState + PathHead =
// LMC_PAYLOAD_U4 x 16 B) as chunk-major SoA, payload[chunk * cap + slot], so a warp moves it with
// fully coalesced 512-byte accesses and a shade kernel touches the 5 KB chain record only for the
// one path vertex it works on.  A chain has at most one pending ray, so the two queues of a step
// kind (light / camera subpath) share one buffer of `cap` = numChains slots and grow towards each
// other from its two ends.
struct alignas(16) Payload { TraceState ts; PathHead ph; };     // working form inside a kernel
// Wire form in the queues: only what the NEXT stage reads.  A stage overwrites the intersection record and
// wi of its subpath state before using them, lastBsdfPdf is never read, and of the PathHead only the
// counters, the light id and the screen position are consulted after the first stage -- so 160 B travel
// instead of 320.  The rest of the head is written to the proposal's record by the first stage (fields a
// later stage sets -- envLight, lensVertexPos -- go straight to the record), and the full light-subpath
// state a connection needs at the last camera vertex is parked in the record (ChainState::lpsFull) by the
// last light-subpath stage.  The ray stays at words 6..13 (k_trace).
struct alignas(16) PayloadLite {
    int stage, depth, offsetId, nLightStates;
    float ndSaved; int ndAvail; float minT, maxT;
    float orgx, orgy, orgz, dirx;
    float diry, dirz; uint32_t rngLo, rngHi;
    uint32_t rngEpoch; int curIdx; float cAccPrev, cAccThis;
    float cThrX, cThrY, cThrZ, cSsJ;
    float lAccPrev, lAccThis, lThrX, lThrY;
    float lThrZ, lSsJ; int lgtLight, lgtPrim;
    int nCam, nLgt, camDepth, lgtDepth;
    float screenX, screenY; int pad0, pad1;
};
#define LMC_PAYLOAD_U4 ((int)(sizeof(PayloadLite) / 16))
#define LMC_ENV_SENTINEL (-2)
__device__ __forceinline__ void payload_pack(const Payload &p, PayloadLite &l) {
    const TraceState &t = p.ts;
    l.stage = t.stage; l.depth = t.depth; l.offsetId = t.offsetId; l.nLightStates = t.nLightStates;
    l.ndSaved = t.ndSaved; l.ndAvail = t.ndAvail; l.minT = t.minT; l.maxT = t.maxT;
    l.orgx = t.ray.org.x; l.orgy = t.ray.org.y; l.orgz = t.ray.org.z; l.dirx = t.ray.dir.x;
    l.diry = t.ray.dir.y; l.dirz = t.ray.dir.z; l.rngLo = t.rngLo; l.rngHi = t.rngHi;
    l.rngEpoch = t.rngEpoch; l.curIdx = t.curIdx; l.cAccPrev = t.cps.accMISWPrev; l.cAccThis = t.cps.accMISWThis;
    l.cThrX = t.cps.throughput.x; l.cThrY = t.cps.throughput.y; l.cThrZ = t.cps.throughput.z; l.cSsJ = t.cps.ssJacobian;
    l.lAccPrev = t.lps.accMISWPrev; l.lAccThis = t.lps.accMISWThis; l.lThrX = t.lps.throughput.x; l.lThrY = t.lps.throughput.y;
    l.lThrZ = t.lps.throughput.z; l.lSsJ = t.lps.ssJacobian; l.lgtLight = p.ph.lgtLight; l.lgtPrim = p.ph.lgtPrim;
    l.nCam = p.ph.nCam; l.nLgt = p.ph.nLgt; l.camDepth = p.ph.camDepth; l.lgtDepth = p.ph.lgtDepth;
    l.screenX = p.ph.screenPos.x; l.screenY = p.ph.screenPos.y; l.pad0 = 0; l.pad1 = 0;
}
__device__ __forceinline__ void payload_unpack(const PayloadLite &l, Payload &p) {
    memset(&p, 0, sizeof(p));
    TraceState &t = p.ts;
    t.stage = l.stage; t.depth = l.depth; t.offsetId = l.offsetId; t.nLightStates = l.nLightStates;
    t.ndSaved = l.ndSaved; t.ndAvail = l.ndAvail; t.minT = l.minT; t.maxT = l.maxT;
    t.ray.org = mk3(l.orgx, l.orgy, l.orgz); t.ray.dir = mk3(l.dirx, l.diry, l.dirz);
    t.rngLo = l.rngLo; t.rngHi = l.rngHi; t.rngEpoch = l.rngEpoch; t.curIdx = l.curIdx;
    t.cps.accMISWPrev = l.cAccPrev; t.cps.accMISWThis = l.cAccThis; t.cps.throughput = mk3(l.cThrX, l.cThrY, l.cThrZ); t.cps.ssJacobian = l.cSsJ;
    t.lps.accMISWPrev = l.lAccPrev; t.lps.accMISWThis = l.lAccThis; t.lps.throughput = mk3(l.lThrX, l.lThrY, l.lThrZ); t.lps.ssJacobian = l.lSsJ;
    p.ph.lgtLight = l.lgtLight; p.ph.lgtPrim = l.lgtPrim;
    p.ph.nCam = l.nCam; p.ph.nLgt = l.nLgt; p.ph.camDepth = l.camDepth; p.ph.lgtDepth = l.lgtDepth;
    p.ph.screenPos = mk2(l.screenX, l.screenY);
}
struct RayQueue {
    int *chain;       // chain slot of entry s
    uint4 *payload;   // [LMC_PAYLOAD_U4][cap]
    float4 *hit;      // x = tid (int bits), y = t, z = u, w = v     (written by k_trace)
    int *count;
    int base, dirn;   // entry t lives in slot base + dirn * t
    int cap;
};
struct ShadowQueue {
    float4 *org;      // xyz, w = dist
    float4 *dir;
    int **flag;       // candidate flag to resolve
    int *chain;       // small steps: the chain whose (only) candidate this is -- its POST part runs right after the
                      // flag is resolved; -1 for large-step candidates (k_prop_post_large handles those)
    int *count;
    int cap;
};
// deferred ConnectVertex work items: (chain, camDepth, lgtDepth, slot)
struct ConnQueue {
    int4 *item;
    int *count;
    int cap;
};
struct WaveQueues {
    RayQueue q[2][4]; // [set][stage - 1]
    ShadowQueue sh;
    ConnQueue cq;
    RunParams rp;          // for the POST part of small steps (filled per lmc_run_chains call)
    SortList propGrad;
};

struct DevShadowSink {
    static const bool kDeferConnections = true;
    ShadowQueue sh;
    ConnQueue cq;
    CamSnap *snaps;      // GenWork::snap of this chain
    int chain, curIdx;
    int postChain;       // = chain for small-step candidates, -1 for large steps
    const Scene *sc;
    // ConnectVertex of a large step: snapshot the camera vertex once, queue one work item per light vertex;
    // k_connect evaluates the pair into the reserved slot.  Queue full: evaluate here.
    __device__ __forceinline__ void emit_connection(const Scene &scn, int camDepth, int lgtDepth, int slot, const BidirPathState *ls,
                                                    const SurfaceVertex *lgtVerts, const BidirPathState &cps, const SurfaceVertex &camVertex,
                                                    V2 screenPos, SubpathContrib *c, int *flag) {
        const int pos = warp_agg_inc(cq.count);
        if (pos < cq.cap) {
            CamSnap &sn = snaps[camDepth];
            if (sn.pad[0] != camDepth + 1) {       // first pair of this camera vertex (pad[0] is reset by k_prop_start)
                sn.cps = cps; sn.tid = camVertex.tid; sn.st = camVertex.st; sn.screenPos = screenPos; sn.curIdx = curIdx;
                sn.pad[0] = camDepth + 1;
            }
            cq.item[pos] = make_int4(chain, camDepth, lgtDepth, slot);
        } else {
            SlotList<DevShadowSink> sl; sl.c = c; sl.flag = flag; sl.sink = this; sl.pend = false;
            const SurfaceVertex lv = lgtVerts[lgtDepth];
            connect_vertex(scn, camDepth, lgtDepth, ls[lgtDepth], lv, cps, camVertex, screenPos, sl);
        }
    }
    __device__ __forceinline__ void emit(const Ray &ray, float dist, int, int *flag) {
        const int pos = warp_agg_inc(sh.count);
        if (pos < sh.cap) {
            sh.org[pos] = make_float4(ray.org.x, ray.org.y, ray.org.z, dist);
            sh.dir[pos] = make_float4(ray.dir.x, ray.dir.y, ray.dir.z, 0.0f);
            sh.flag[pos] = flag;
            sh.chain[pos] = postChain;
        } else {
            *flag = cand_resolve(*flag, scene_occluded(*sc, ray, dist));    // queue full: resolve on the spot
        }
    }
};

__device__ __forceinline__ void payload_load(const RayQueue &q, int slot, Payload &p) {
    PayloadLite l;
    uint4 *d = reinterpret_cast<uint4 *>(&l);
#pragma unroll
    for (int k = 0; k < LMC_PAYLOAD_U4; k++) d[k] = q.payload[(size_t)k * q.cap + slot];
    payload_unpack(l, p);
}
__device__ __forceinline__ void ray_push(const RayQueue &q, bool pred, int chain, const Payload &p) {
    const unsigned mask = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(q.count, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    const int slot = q.base + q.dirn * (base + __popc(mask & ((1u << lane) - 1u)));
    q.chain[slot] = chain;
    PayloadLite l;
    payload_pack(p, l);
    const uint4 *s = reinterpret_cast<const uint4 *>(&l);
#pragma unroll
    for (int k = 0; k < LMC_PAYLOAD_U4; k++) q.payload[(size_t)k * q.cap + slot] = s[k];
}
template <class T>
__device__ __forceinline__ void copy_u4(T &dst, const T &src) {
    static_assert(sizeof(T) % 16 == 0 && alignof(T) >= 16, "16-byte records only");
    uint4 *d = reinterpret_cast<uint4 *>(&dst);
    const uint4 *s = reinterpret_cast<const uint4 *>(&src);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = s[k];
}
// the two offsets a perturbation stage consumes, fetched up front
struct OffPair {
    float v0, v1; int base;
    __device__ __forceinline__ float operator[](int i) const { return (i == base) ? v0 : v1; }
};
__device__ __forceinline__ void rng_from_payload(Rng &rng, uint32_t *tab, const Scene &sc, int globalChainId, const TraceState &ts) {
    rng.tab = tab; rng.stride = 1;
    const uint64_t seed = (uint64_t)(long long)(globalChainId + sc.opt.seedOffset);
    rng_restore_lazy(rng, seed, ((uint64_t)ts.rngHi << 32) | (uint64_t)ts.rngLo, ts.rngEpoch);
}
__device__ __forceinline__ void rng_to_payload(const Rng &rng, TraceState &ts) {
    ts.rngLo = (uint32_t)rng.state; ts.rngHi = (uint32_t)(rng.state >> 32); ts.rngEpoch = rng.epoch;
}

#ifndef LMC_SHADE_BLOCK
#define LMC_SHADE_BLOCK 256        // 2 blocks of 8 warps per SM; the warps of a block re-align at every queue item
#endif
#ifndef LMC_SHADE_MINB
#define LMC_SHADE_MINB 2
#endif
#ifndef LMC_START_MINB
#define LMC_START_MINB LMC_SHADE_MINB
#endif
template <int MAXD> struct GenWorkT { typedef GenWork<MAXD, Limits<MAXD>::MAXC> type; };

// first stage of every proposal: PRE part of the mutation + the statements up to the first ray
template <int MAXD, int LARGE>
__global__ void __launch_bounds__(LMC_SHADE_BLOCK, LMC_START_MINB) k_prop_start(const __grid_constant__ Scene sc, int chainBase, ChainRec<MAXD> *states,
                                                                 typename GenWorkT<MAXD>::type *genWork, const int *list, const int *countp,
                                                                 WaveQueues wq, H2mcSide *sides) {
    const int count = *countp;
    const int stride = gridDim.x * blockDim.x;
    for (int base = blockIdx.x * blockDim.x; base < count; base += stride) {
        const int t = base + threadIdx.x;
        bool more = false; int i = -1;
        Payload p;
        if (t < count) {
            i = list[t];
            uint32_t tab[64];
            ChainState<MAXD> &cs = states[i].cs;
            Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
            const int curIdx = cs.curIdx;
            MarkovState<MAXD> &cur = cs.st[curIdx], &prop = cs.st[curIdx ^ 1];
            memset(&p.ts, 0, sizeof(p.ts));
            p.ts.curIdx = curIdx;
            if (LARGE) {
                cs.ch.lastMutationType = MUT_LARGE;             // propose_pre_large on the payload's head
                copy_u4<PathHead>(p.ph, prop.path);
                path_clear(p.ph);
                genWork[i].n = 0;
                for (int d = 0; d < MAXD; d++) genWork[i].snap[d].pad[0] = 0;
                more = gen_stage_begin(sc, p.ph, p.ts, genWork[i].ls, rng);
            } else {
                propose_pre_small<MAXD, false>(sc, cur, prop, cs.ch, rng, cs.ss, sides ? sides + i : nullptr, curIdx);
                cs.pc.n = 0;
                copy_u4<PathHead>(p.ph, cur.path);
                const float *offset = cs.ss.offset;
                more = perturb_stage_begin(sc, offset, p.ph, p.ts, rng);
            }
            rng_to_payload(rng, p.ts);
            copy_u4<PathHead>(prop.path, p.ph);        // the whole head now; later stages update single fields
            if (!more) { rng_close(rng, cs); if (!LARGE) post_small_now(sc, wq, cs, i); }
        }
        if (LARGE) {
            ray_push(wq.q[0][TS_G_LGT - 1], more, i, p);
        } else {
            ray_push(wq.q[0][TS_P_LGT - 1], more && p.ts.stage == TS_P_LGT, i, p);
            ray_push(wq.q[0][TS_P_CAM - 1], more && p.ts.stage == TS_P_CAM, i, p);
        }
    }
}

// POST part of a small step (mutation.h propose_post_small) + its proposal-gradient key.  Runs as soon as the
// chain's candidate contribution is resolved: in the stage that ends the proposal when no shadow ray is
// pending, else in k_shadow right after the ray has been traced.
template <int MAXD>
__device__ __forceinline__ void post_small_now(const Scene &sc, const WaveQueues &wq, ChainState<MAXD> &cs, int i) {
    MarkovState<MAXD> &cur = cs.st[cs.curIdx], &prop = cs.st[cs.curIdx ^ 1];
    cs.pc.n = deferred_compact(cs.pc.c, cs.pc.flag, cs.pc.n);
    propose_post_small(sc, wq.rp, cur, prop, cs.ss, cs.pc);
    // key = the class the evaluator will see (path.camDepth / path.lgtDepth): blocks must be class-pure
    sort_key_set(wq.propGrad, i, cs.ss.needPropGrad ? class_key(prop.path.camDepth, prop.path.lgtDepth, 0) : -1);
}
template <int MAXD>
__device__ __forceinline__ bool small_candidate_pending(const ChainState<MAXD> &cs) {
    bool pending = false;
    for (int k = 0; k < cs.pc.n; k++) pending = pending || ((cs.pc.flag[k] & CAND_PENDING) != 0);
    return pending;
}

// One stage of one proposal: the statements between two ray queries for the chain `i` whose pending ray
// of stage STAGE was answered with `hit`.  Returns true when the proposal wants another ray (p.ts.stage
// tells of which kind); on false it has left the wavefront and its head / RNG state are back in the record.
template <int MAXD, int STAGE>
__device__ __forceinline__ bool shade_entry(const Scene &sc, int chainBase, ChainRec<MAXD> *states, typename GenWorkT<MAXD>::type *genWork,
                                            const WaveQueues &wq, int i, Payload &p, const Hit &hit) {
    ChainState<MAXD> &cs = states[i].cs;
    MarkovState<MAXD> &cur = cs.st[p.ts.curIdx], &prop = cs.st[p.ts.curIdx ^ 1];
    uint32_t tab[64];
    Rng rng; rng_from_payload(rng, tab, sc, chainBase + i, p.ts);
    DevShadowSink sink; sink.sh = wq.sh; sink.sc = &sc; sink.cq = wq.cq; sink.chain = i; sink.curIdx = p.ts.curIdx;
    sink.snaps = genWork ? genWork[i].snap : nullptr;
    sink.postChain = (STAGE == TS_P_LGT || STAGE == TS_P_CAM) ? i : -1;
    DeferredList<DevShadowSink> dl;
    SurfaceVertex sv;
    bool more;
    // head fields a stage may SET are detected through sentinels and written straight to the record
    p.ph.envLight = LMC_ENV_SENTINEL;
    p.ph.lensVertexPos.x = __int_as_float(0x7fc00000);
    if (STAGE == TS_P_LGT || STAGE == TS_P_CAM) {
        OffPair off; off.base = p.ts.offsetId;
        off.v0 = cs.ss.offset[off.base]; off.v1 = cs.ss.offset[off.base + 1];
        dl.bind(cs.pc.c, cs.pc.flag, &cs.pc.n, 2, &sink);
        const int d = p.ts.depth;
        if (STAGE == TS_P_LGT) {
            copy_u4<SurfaceVertex>(sv, cur.path.lgt[d]);
            more = perturb_stage_light(sc, off, p.ph, sv, p.ts, dl, rng, hit);
            copy_u4<SurfaceVertex>(prop.path.lgt[d], sv);
            // light subpath complete: park its state for the connection at the end of the camera subpath
            if (more && p.ts.stage == TS_P_CAM) cs.lpsFull.s = p.ts.lps;
        } else {
            copy_u4<SurfaceVertex>(sv, cur.path.cam[d]);
            // the connection at the last camera vertex needs the whole light-subpath state
            if (p.ph.lgtDepth > 1 && d == p.ph.nCam - 1) p.ts.lps = cs.lpsFull.s;
            more = perturb_stage_camera(sc, off, p.ph, sv, prop.path.lgt, p.ts, dl, rng, hit);
            copy_u4<SurfaceVertex>(prop.path.cam[d], sv);
        }
    } else {
        typename GenWorkT<MAXD>::type &gw = genWork[i];
        dl.bind(gw.c, gw.flag, &gw.n, Limits<MAXD>::MAXC, &sink);
        const int minDepth = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
        if (STAGE == TS_G_LGT) {
            const int d = p.ph.nLgt;
            more = gen_stage_light(sc, minDepth, sc.opt.maxDepth, p.ph, sv, p.ts, gw.ls, dl, rng, hit);
            copy_u4<SurfaceVertex>(prop.path.lgt[d], sv);
        } else {
            const int d = p.ph.nCam;
            more = gen_stage_camera(sc, minDepth, sc.opt.maxDepth, p.ph, sv, prop.path.lgt, p.ts, gw.ls, dl, rng, hit);
            copy_u4<SurfaceVertex>(prop.path.cam[d], sv);
        }
    }
    rng_to_payload(rng, p.ts);
    if (p.ph.envLight != LMC_ENV_SENTINEL) { prop.path.envLight = p.ph.envLight; prop.path.envPrim = p.ph.envPrim; }
    if (p.ph.lensVertexPos.x == p.ph.lensVertexPos.x) prop.path.lensVertexPos = p.ph.lensVertexPos;
    if (!more) {      // leaving the wavefront: counters, screen position and RNG state back to the record
        prop.path.nCam = p.ph.nCam; prop.path.nLgt = p.ph.nLgt; prop.path.screenPos = p.ph.screenPos;
        rng_close(rng, cs);
        if ((STAGE == TS_P_LGT || STAGE == TS_P_CAM) && !small_candidate_pending(cs)) post_small_now(sc, wq, cs, i);
    }
    return more;
}

template <int MAXD, int STAGE>
__global__ void __launch_bounds__(LMC_SHADE_BLOCK, LMC_SHADE_MINB) k_shade(const __grid_constant__ Scene sc, int chainBase, ChainRec<MAXD> *states,
                                                            typename GenWorkT<MAXD>::type *genWork, WaveQueues wq, int curSet) {
    const RayQueue &qi = wq.q[curSet][STAGE - 1];
    const int count = *qi.count;
    const int stride = gridDim.x * blockDim.x;
    const int nextSet = curSet ^ 1;
    for (int base = blockIdx.x * blockDim.x; base < count; base += stride) {
#ifndef LMC_SHADE_NOSYNC
#ifdef LMC_SHADE_SYNC_EVERY
        if (((base / stride) % LMC_SHADE_SYNC_EVERY) == 0)
#endif
        __syncthreads();          // keep the block's warps in the same stretch of code (instruction fetch)
#endif
        const int t = base + threadIdx.x;
        bool more = false; int i = -1;
        Payload p;
        if (t < count) {
            const int slot = qi.base + qi.dirn * t;
            i = qi.chain[slot];
            payload_load(qi, slot, p);
            const float4 h4 = qi.hit[slot];
            Hit hit; hit.tid = __float_as_int(h4.x); hit.t = h4.y; hit.u = h4.z; hit.v = h4.w;
            more = shade_entry<MAXD, STAGE>(sc, chainBase, states, genWork, wq, i, p, hit);
        }
        if (STAGE == TS_P_LGT) {
            ray_push(wq.q[nextSet][TS_P_LGT - 1], more && p.ts.stage == TS_P_LGT, i, p);
            ray_push(wq.q[nextSet][TS_P_CAM - 1], more && p.ts.stage == TS_P_CAM, i, p);
        } else if (STAGE == TS_P_CAM) {
            ray_push(wq.q[nextSet][TS_P_CAM - 1], more, i, p);
        } else if (STAGE == TS_G_LGT) {
            ray_push(wq.q[nextSet][TS_G_LGT - 1], more && p.ts.stage == TS_G_LGT, i, p);
            ray_push(wq.q[nextSet][TS_G_CAM - 1], more && p.ts.stage == TS_G_CAM, i, p);
        } else {
            ray_push(wq.q[nextSet][TS_G_CAM - 1], more, i, p);
        }
    }
}

// Tail of the wavefront: after LMC_FULL_WAVES waves only the few long paths are still alive (< 5 % of the
// rays of an iteration, spread over up to 2 * maxDepth - 6 more waves of nearly empty launches).  This
// kernel finishes them in one launch, one thread per proposal looping "closest hit, next stage" to the end.
#ifndef LMC_FULL_WAVES
#define LMC_FULL_WAVES 6          // before the calibration iteration
#endif
#ifndef LMC_TAIL_DIV
#define LMC_TAIL_DIV 16           // measured: 12..32 are equivalent on torus L8 (6..8 waves) and door L12 (16..18)
#endif
template <int MAXD>
__global__ void __launch_bounds__(128) k_shade_tail(const __grid_constant__ Scene sc, int chainBase, ChainRec<MAXD> *states,
                                                    typename GenWorkT<MAXD>::type *genWork, WaveQueues wq, int curSet) {
    const RayQueue *q = wq.q[curSet];
    const int c0 = *q[0].count, c1 = c0 + *q[1].count, c2 = c1 + *q[2].count, c3 = c2 + *q[3].count;
    const int stride = gridDim.x * blockDim.x;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < c3; idx += stride) {
        const int k = (idx >= c0) + (idx >= c1) + (idx >= c2);
        const int pos = idx - (k == 0 ? 0 : (k == 1 ? c0 : (k == 2 ? c1 : c2)));
        const int slot = q[k].base + q[k].dirn * pos;
        const int i = q[k].chain[slot];
        Payload p;
        payload_load(q[k], slot, p);
        bool more = true;
        while (more) {
            const Hit hit = bvh_traverse<false>(sc, p.ts.ray, p.ts.minT, p.ts.maxT);
            switch (p.ts.stage) {
                case TS_P_LGT: more = shade_entry<MAXD, TS_P_LGT>(sc, chainBase, states, genWork, wq, i, p, hit); break;
                case TS_P_CAM: more = shade_entry<MAXD, TS_P_CAM>(sc, chainBase, states, genWork, wq, i, p, hit); break;
                case TS_G_LGT: more = shade_entry<MAXD, TS_G_LGT>(sc, chainBase, states, genWork, wq, i, p, hit); break;
                default:       more = shade_entry<MAXD, TS_G_CAM>(sc, chainBase, states, genWork, wq, i, p, hit); break;
            }
        }
    }
}

// closest hit for the four ray queues of a wave (persistent warps, trace_kernels.cuh); the ray sits in
// payload words 6..13
struct ClosestSrc {
    const RayQueue *q;     // the 4 queues of the set
    int c0, c1, c2, c3;    // running totals of their counts
    __device__ __forceinline__ int total() const { return c3; }
    __device__ __forceinline__ void locate(int idx, int &k, int &slot) const {
        k = (idx >= c0) + (idx >= c1) + (idx >= c2);
        const int pos = idx - (k == 0 ? 0 : (k == 1 ? c0 : (k == 2 ? c1 : c2)));
        slot = q[k].base + q[k].dirn * pos;
    }
    __device__ __forceinline__ void load(int idx, Ray &ray, float &minT, float &maxT) const {
        int k, slot; locate(idx, k, slot);
        const RayQueue &qq = q[k];
        const uint4 a = qq.payload[(size_t)1 * qq.cap + slot], b = qq.payload[(size_t)2 * qq.cap + slot], c = qq.payload[(size_t)3 * qq.cap + slot];
        ray.org = mk3(__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z));
        ray.dir = mk3(__uint_as_float(b.w), __uint_as_float(c.x), __uint_as_float(c.y));
        minT = __uint_as_float(a.z); maxT = __uint_as_float(a.w);
    }
    __device__ __forceinline__ void store(int idx, const Hit &h) const {
        int k, slot; locate(idx, k, slot);
        q[k].hit[slot] = make_float4(__int_as_float(h.tid), h.t, h.u, h.v);
    }
};
// any hit for the queued connection segments: Occluded(scene, ray, dist), src/scene.cpp:128-149
template <int MAXD>
struct ShadowSrc {
    const Scene *sc; const WaveQueues *wq; ChainRec<MAXD> *states; int n;
    __device__ __forceinline__ int total() const { return n; }
    __device__ __forceinline__ void load(int idx, Ray &ray, float &minT, float &maxT) const {
        const float4 o = wq->sh.org[idx], d = wq->sh.dir[idx];
        ray.org = mk3(o.x, o.y, o.z); ray.dir = mk3(d.x, d.y, d.z);
        minT = LMC_ISECT_EPS;
        maxT = (o.w == dm_inf()) ? dm_inf() : (1.0f - LMC_SHADOW_EPS) * o.w;
    }
    __device__ __forceinline__ void store(int idx, const Hit &h) const {
        int *f = wq->sh.flag[idx];
        *f = cand_resolve(*f, h.tid >= 0);
        const int chain = wq->sh.chain[idx];
        if (chain >= 0) post_small_now(*sc, *wq, states[chain].cs, chain);      // a small step has exactly one pending candidate
    }
};
#ifndef LMC_TRACE_MINB
#define LMC_TRACE_MINB 2
#endif
static __global__ void __launch_bounds__(LMC_TRACE_BLOCK, LMC_TRACE_MINB) k_trace(const __grid_constant__ Scene sc, WaveQueues wq, int curSet, int *cursor) {
    __shared__ __align__(128) BvhNode top[LMC_TOP_NODES];
    __shared__ uint64_t bar;
    const int topCount = sc.numNodes < LMC_TOP_NODES ? sc.numNodes : LMC_TOP_NODES;
    ClosestSrc src; src.q = wq.q[curSet];
    src.c0 = *wq.q[curSet][0].count; src.c1 = src.c0 + *wq.q[curSet][1].count; src.c2 = src.c1 + *wq.q[curSet][2].count;
    src.c3 = src.c2 + *wq.q[curSet][3].count;
    if ((long long)blockIdx.x * LMC_TRACE_BLOCK >= (long long)src.c3) return;     // more blocks than rays: leave before staging
    tma_stage_nodes(top, sc.nodes, topCount, &bar);
    trace_persistent<false>(sc, top, topCount, src, cursor);
}
template <int MAXD>
__global__ void __launch_bounds__(LMC_TRACE_BLOCK, LMC_TRACE_MINB) k_shadow(const __grid_constant__ Scene sc, ChainRec<MAXD> *states,
                                                                          const __grid_constant__ WaveQueues wq, int *cursor) {
    __shared__ __align__(128) BvhNode top[LMC_TOP_NODES];
    __shared__ uint64_t bar;
    const int topCount = sc.numNodes < LMC_TOP_NODES ? sc.numNodes : LMC_TOP_NODES;
    ShadowSrc<MAXD> src; src.sc = &sc; src.wq = &wq; src.states = states;
    src.n = *wq.sh.count; if (src.n > wq.sh.cap) src.n = wq.sh.cap;
    if ((long long)blockIdx.x * LMC_TRACE_BLOCK >= (long long)src.n) return;
    tma_stage_nodes(top, sc.nodes, topCount, &bar);
    trace_persistent<true>(sc, top, topCount, src, cursor);
}

// deferred ConnectVertex of the large steps: one thread per (camera vertex, light vertex) pair
template <int MAXD>
__global__ void __launch_bounds__(128) k_connect(const __grid_constant__ Scene sc, ChainRec<MAXD> *states, typename GenWorkT<MAXD>::type *genWork,
                                                 WaveQueues wq) {
    int n = *wq.cq.count; if (n > wq.cq.cap) n = wq.cq.cap;
    const int stride = gridDim.x * blockDim.x;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int4 it = wq.cq.item[idx];
        typename GenWorkT<MAXD>::type &gw = genWork[it.x];
        const CamSnap &sn = gw.snap[it.y];
        DevShadowSink sink; sink.sh = wq.sh; sink.sc = &sc; sink.cq = wq.cq; sink.chain = it.x; sink.curIdx = sn.curIdx; sink.snaps = gw.snap;
        sink.postChain = -1;         // large-step candidate: k_prop_post_large finishes the proposal
        SlotList<DevShadowSink> sl; sl.c = gw.c + it.w; sl.flag = gw.flag + it.w; sl.sink = &sink; sl.pend = false;
        const SurfaceVertex lv = states[it.x].cs.st[sn.curIdx ^ 1].path.lgt[it.z];
        SurfaceVertex cv; cv.tid = sn.tid; cv.st = sn.st;
        connect_vertex(sc, it.y, it.z, gw.ls[it.z], lv, sn.cps, cv, sn.screenPos, sl);
    }
}

// POST part of the large steps once every candidate is resolved (mutation.h propose_post_large): candidate
// compaction, contribution choice, acceptance.  (Small steps: post_small_now, above.)
template <int MAXD>
__global__ void __launch_bounds__(LMC_CHAIN_BLOCK) k_prop_post_large(const __grid_constant__ Scene sc, RunParams rp, int chainBase, ChainRec<MAXD> *states,
                                                                      typename GenWorkT<MAXD>::type *genWork, const int *list, const int *countp,
                                                                      WaveLists wl) {
    const int count = *countp;
    const int stride = gridDim.x * blockDim.x;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const int i = list[t];
        ChainState<MAXD> &cs = states[i].cs;
        MarkovState<MAXD> &cur = cs.st[cs.curIdx], &prop = cs.st[cs.curIdx ^ 1];
        uint32_t tab[64];
        Rng rng; rng_open(rng, tab, sc, chainBase + i, cs);
        typename GenWorkT<MAXD>::type &gw = genWork[i];
        gw.n = deferred_compact(gw.c, gw.flag, gw.n);
        propose_post_large(rp, cur, prop, cs.ch, rng, cs.ss, gw);
        rng_close(rng, cs);
        sort_key_set(wl.propGrad, i, -1);          // a large step evaluates no proposal gradient
    }
}

// ---- MLTInit on the device (src/mlt.h:41-106): the init paths ---------------------------------
// Logical thread t = one CUDA thread: RNG(t + seedOffset), its share of the init samples generated
// back to back with GeneratePathBidir exactly like the host restatement (host/mlt_init.h).  Pass 1
// (EMIT = 0) counts the contributions of every logical thread, the host turns the counts into
// offsets, pass 2 (EMIT = 1) regenerates the same paths and writes their lsScores in
// (logical thread, sample, contribution) order.  The fp32 running sums over that list (CDF,
// normalisation) stay sequential on the host so that both forms agree bit for bit.
template <int MAXD, int EMIT>
__global__ void __launch_bounds__(128) k_mlt_init_paths(const __grid_constant__ Scene sc, long long numInitSamples, int logicalThreads,
                                                        int tBegin, int tEnd, int *counts, const long long *offsets, float *scores) {
    // logical threads [tBegin, tEnd) of logicalThreads: a shard of the init pass (one per GPU); counts / offsets are
    // indexed relative to tBegin
    const int tl = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = tBegin + tl;
    if (t >= tEnd) return;
    uint32_t tab[64];
    Rng rng; rng.tab = tab; rng.stride = 1;
    rng_seed_lazy(rng, (uint64_t)(long long)(t + sc.opt.seedOffset));
    const long long perT = numInitSamples / logicalThreads, extra = numInitSamples % logicalThreads;
    const long long n = perT + ((t < extra) ? 1 : 0);
    const int minPathLength = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
    Path<MAXD> path;
    ContribList<Limits<MAXD>::MAXC> contribs;
    int count = 0;
    long long pos = EMIT ? offsets[tl] : 0;
    for (long long s = 0; s < n; s++) {
        contribs.clear();
        path_clear(path);
        generate_path_bidir(sc, minPathLength, sc.opt.maxDepth, path, contribs, rng);
        if (EMIT) for (int i = 0; i < contribs.n; i++) scores[pos++] = contribs.c[i].lsScore;
        count += contribs.n;
    }
    if (!EMIT) counts[tl] = count;
}

template <int MAXD>
__global__ void k_chain_stats(const ChainRec<MAXD> *states, int n, unsigned long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[13];
    for (int k = 0; k < 13; k++) v[k] = 0ULL;
    if (i < n) {
        const ChainState<MAXD> &cs = states[i].cs;
        for (int k = 0; k < 4; k++) { v[k] = cs.nPropose[k]; v[4 + k] = cs.nAccept[k]; }
        v[8] = cs.gradStats[0]; v[9] = cs.gradStats[1]; v[10] = (unsigned long long)cs.ch.outlierResets;
        v[11] = (unsigned long long)cs.ch.cacheQueries; v[12] = (unsigned long long)cs.ch.cacheHits;
    }
    for (int k = 0; k < 13; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, x);
    }
}


// ---- global cache: commit of one iteration's push requests in chain order (core/chain.h cache_commit_host) --------
// k_cache_count   per block and slot: number of chains asking for a push            -> blockCounts[slot][block]
// k_cache_scan    one block: exclusive scan over the blocks, starting at the slot's fill level; new fill level and
//                 ready flag (slot full); `active` = any request at all this iteration
// k_cache_write   stable rank of a request inside its block (warp ballots) + block offset = its entry index;
//                 requests that land beyond PSS_MAX_SIZE are dropped; the request flag is cleared
// All three return at once when every slot is ready (the steady state of a cache run).
#define LMC_CACHE_BLOCK 256
__device__ __forceinline__ bool cache_all_ready(const GlobalCacheView &gc) {
    bool all = true;
    for (int s = 0; s < LMC_CACHE_SLOTS; s++) all = all && gc.ready[s] != 0;
    return all;
}
template <int MAXD>
__global__ void __launch_bounds__(LMC_CACHE_BLOCK) k_cache_count(const __grid_constant__ Scene sc, const ChainRec<MAXD> *states, int n, int *blockCounts) {
    if (cache_all_ready(sc.gc)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = (i < n) ? cache_slot(states[i].cs.ch.pushDim) : -1;
    for (int s = 0; s < LMC_CACHE_SLOTS; s++) {
        const int c = __syncthreads_count(slot == s);
        if (threadIdx.x == 0) blockCounts[s * gridDim.x + blockIdx.x] = c;
    }
}
static __global__ void __launch_bounds__(1024) k_cache_scan(GlobalCacheView gc, int *blockCounts, int numBlocks, int *active) {
    __shared__ int warpSum[32];
    __shared__ int carry;
    if (cache_all_ready(gc)) { if (threadIdx.x == 0) *active = 0; return; }
    int any = 0;
    for (int s = 0; s < LMC_CACHE_SLOTS; s++) {
        if (threadIdx.x == 0) carry = gc.count[s];
        __syncthreads();
        for (int b0 = 0; b0 < numBlocks; b0 += 1024) {
            const int b = b0 + threadIdx.x;
            const int v = (b < numBlocks) ? blockCounts[s * numBlocks + b] : 0;
            int x = v;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
            if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = x;
            __syncthreads();
            if (threadIdx.x < 32) {
                int w = warpSum[threadIdx.x];
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
                warpSum[threadIdx.x] = w;
            }
            __syncthreads();
            const int before = carry + ((threadIdx.x >> 5) ? warpSum[(threadIdx.x >> 5) - 1] : 0) + x - v;
            if (b < numBlocks) blockCounts[s * numBlocks + b] = before;      // exclusive prefix = first entry index of the block
            __syncthreads();
            if (threadIdx.x == 1023) carry = before + v;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            any |= (carry > gc.count[s]);
            gc.count[s] = carry < LMC_CACHE_MAX_SIZE ? carry : LMC_CACHE_MAX_SIZE;
            if (carry >= LMC_CACHE_MAX_SIZE) gc.ready[s] = 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *active = any;
}
template <int MAXD>
__global__ void __launch_bounds__(LMC_CACHE_BLOCK) k_cache_write(const __grid_constant__ Scene sc, ChainRec<MAXD> *states, int n, const int *blockCounts,
                                                                const int *active) {
    if (!*active) return;
    __shared__ int warpCnt[LMC_CACHE_SLOTS][LMC_CACHE_BLOCK / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int dim = (i < n) ? states[i].cs.ch.pushDim : 0;
    const int slot = cache_slot(dim);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int rankInWarp = 0;
    for (int s = 0; s < LMC_CACHE_SLOTS; s++) {
        const unsigned m = __ballot_sync(0xffffffffu, slot == s);
        if (slot == s) rankInWarp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warpCnt[s][warp] = __popc(m);
    }
    __syncthreads();
    if (slot >= 0) {
        int rank = rankInWarp;
        for (int w = 0; w < warp; w++) rank += warpCnt[slot][w];
        const int pos = blockCounts[slot * gridDim.x + blockIdx.x] + rank;
        ChainVars<MAXD> &ch = states[i].cs.ch;
        if (pos < LMC_CACHE_MAX_SIZE) {
            float *e = sc.gc.data + cache_slot_offset(slot) + (size_t)pos * 3 * dim;
            for (int k = 0; k < dim; k++) { e[k] = ch.pss[k]; e[dim + k] = ch.v1[k]; e[2 * dim + k] = ch.v2[k]; }
        }
        ch.pushDim = 0;
    }
}

// k_cache_grid    block s: once slot s is ready (and its last entries are written), bin its entries into the query grid
//                 of scene.h (counts by atomics, serial prefix sum over the 13 824 cells, fill by atomic cursors; the
//                 query does not depend on the order inside a cell).  Runs once per slot and job; otherwise returns at once.
static __global__ void __launch_bounds__(256) k_cache_grid(GlobalCacheView gc) {
    const int s = blockIdx.x;
    if (!gc.grid || !gc.ready[s] || gc.gridReady[s]) return;
    const int dim = 4 + 2 * s;
    int *cellStart = gc.grid + (size_t)s * LMC_CACHE_GRID_INTS, *cursor = cellStart + LMC_CACHE_CELLS + 1, *entry = cursor + LMC_CACHE_CELLS;
    const float *base = gc.data + cache_slot_offset(s);
    for (int c = threadIdx.x; c <= LMC_CACHE_CELLS; c += blockDim.x) cellStart[c] = 0;
    for (int c = threadIdx.x; c < LMC_CACHE_CELLS; c += blockDim.x) cursor[c] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < LMC_CACHE_MAX_SIZE; e += blockDim.x) atomicAdd(&cellStart[cache_cell(base + (size_t)e * 3 * dim) + 1], 1);
    __syncthreads();
    if (threadIdx.x == 0) for (int c = 0; c < LMC_CACHE_CELLS; c++) cellStart[c + 1] += cellStart[c];
    __syncthreads();
    for (int e = threadIdx.x; e < LMC_CACHE_MAX_SIZE; e += blockDim.x) {
        const int c = cache_cell(base + (size_t)e * 3 * dim);
        entry[cellStart[c] + atomicAdd(&cursor[c], 1)] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); gc.gridReady[s] = 1; }
}

// k_cache_prequery   the cache query of every proposal that k_wave_finish is about to look up (MALA step with a
//                 contribution, dimension ready, no reuse: the conditions of core/mutation.h mala_finish_gaussian), answered
//                 by a WARP per query: a thread-per-chain query runs as long as the densest cell neighbourhood among its 32
//                 lanes (the PSS entries cluster: 50 .. 500 candidates), here the 32 lanes test 32 candidates at a time.
//                 In-radius entries go to a per-warp shared list; the owner lane keeps the LMC_CACHE_KNN smallest insertion
//                 indices, averages (cache_average) into chain.v1 / v2 and leaves ChainVars::preq = 2 (hit) or 1 (miss) for
//                 the finish kernel -- bit-identical to the inline query.  A list overflow leaves preq = 0 (inline query).
#define LMC_PREQ_CAP 64
template <int MAXD>
__global__ void __launch_bounds__(256) k_cache_prequery(const __grid_constant__ Scene sc, ChainRec<MAXD> *states, int n) {
    const int DIM = Limits<MAXD>::DIM;
    __shared__ int sIdx[8][LMC_PREQ_CAP];
    __shared__ float sDist[8][LMC_PREQ_CAP];
    __shared__ int sCnt[8];
    if (!sc.gc.grid) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool need = false; int dim = 0, slot = -1;
    float pss[DIM];
    for (int j = 0; j < DIM; j++) pss[j] = 0.0f;
    if (i < n) {
        const ChainState<MAXD> &cs = states[i].cs;
        if (cs.ss.kind == STEP_MALA && cs.ss.hasContrib) {
            const MarkovState<MAXD> &prop = cs.st[cs.curIdx ^ 1];
            if (mala_grad_mode(sc, prop) == 3) {
                dim = path_dimension(prop.path);
                slot = cache_slot(dim);
                get_path_pss(prop.path, pss);
                need = slot >= 0 && sc.gc.gridReady[slot] && !cache_reuse(dim, cs.ch.queried, pss, cs.ch.last_pss);
            }
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, need);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int d = __shfl_sync(0xffffffffu, dim, src), s = __shfl_sync(0xffffffffu, slot, src);
        float q[DIM];
#pragma unroll
        for (int j = 0; j < DIM; j++) q[j] = __shfl_sync(0xffffffffu, pss[j], src);
        const float *base = sc.gc.data + cache_slot_offset(s);
        const int stride = 3 * d;
        const float radius = (float)d * (LMC_CACHE_QUERY_DIST * LMC_CACHE_QUERY_DIST);
        const int *cellStart = sc.gc.grid + (size_t)s * LMC_CACHE_GRID_INTS;
        const int *entry = cellStart + 2 * LMC_CACHE_CELLS + 1;
        const int c0 = cache_cell_coord(q[0]), c1 = cache_cell_coord(q[1]), c2 = cache_cell_coord(q[2]);
        const int G = LMC_CACHE_GRID;
        if (lane == 0) sCnt[warp] = 0;
        __syncwarp();
        for (int a = (c0 > 0 ? c0 - 1 : 0); a <= (c0 < G - 1 ? c0 + 1 : G - 1); a++)
            for (int b = (c1 > 0 ? c1 - 1 : 0); b <= (c1 < G - 1 ? c1 + 1 : G - 1); b++) {
                const int row = (a * G + b) * G;
                const int p0 = cellStart[row + (c2 > 0 ? c2 - 1 : 0)], p1 = cellStart[row + (c2 < G - 1 ? c2 + 1 : G - 1) + 1];
                for (int pb = p0; pb < p1; pb += 32) {
                    const int p = pb + lane;
                    bool hit = false; int e = 0; float dd = 0.0f;
                    if (p < p1) {
                        e = entry[p];
                        const float *x = base + (size_t)e * stride;
#pragma unroll
                        for (int j = 0; j < DIM; j++) if (j < d) { const float t = q[j] - x[j]; dd += t * t; }
                        hit = dd < radius;
                    }
                    const unsigned hm = __ballot_sync(0xffffffffu, hit);
                    if (hm) {
                        const int at = sCnt[warp] + __popc(hm & ((1u << lane) - 1u));
                        if (hit && at < LMC_PREQ_CAP) { sIdx[warp][at] = e; sDist[warp][at] = dd; }
                        __syncwarp();
                        if (lane == 0) sCnt[warp] += __popc(hm);
                        __syncwarp();
                    }
                }
            }
        if (lane == src) {
            const int cnt = sCnt[warp];
            ChainVars<MAXD> &ch = states[i].cs.ch;
            if (cnt == 0) ch.preq = 1;
            else if (cnt <= LMC_PREQ_CAP) {
                int idx[LMC_CACHE_KNN]; float dist[LMC_CACHE_KNN]; int found = 0;
                for (int k = 0; k < cnt; k++) {              // the KNN smallest insertion indices, ascending
                    const int e = sIdx[warp][k]; const float dd = sDist[warp][k];
                    int m = found < LMC_CACHE_KNN ? found : LMC_CACHE_KNN - 1;
                    if (found == LMC_CACHE_KNN && e > idx[m]) continue;
                    while (m > 0 && idx[m - 1] > e) { idx[m] = idx[m - 1]; dist[m] = dist[m - 1]; m--; }
                    idx[m] = e; dist[m] = dd;
                    if (found < LMC_CACHE_KNN) found++;
                }
                cache_average(base, d, found, idx, dist, ch.v1, ch.v2);
                ch.preq = 2;
            }
        }
        __syncwarp();
    }
}

// launchers (defined by LMC_INSTANTIATE_CHAIN in chain_inst_*.cu)
#ifndef LMC_WAVEFRONT_MIN_CHAINS
#define LMC_WAVEFRONT_MIN_CHAINS 393216     // measured crossover on B200 (torus, maxdepth 8): 2^18 -> monolithic, 2^19 -> wavefront
#endif
#define LMC_NCOUNTERS 80        // 9 counters + up to 64 wave cursors (maxdepth <= 12: 23 waves + shadow)
struct WaveCfg {
    WaveQueues wq;
    void *genWork;
    int *queueCounts;      // 2 x 4 ray-queue counters, the shadow counter, then one traversal cursor per wave (LMC_NCOUNTERS ints)
    H2mcSide *padSide;     // scratch Hessian for the padding threads of the H2MC gradient kernel
    int wavefront;         // 1: per-vertex wavefront proposal; 0: monolithic k_wave_propose; -1: by chain count
    int smCount;
    // Number of full waves before the tail kernel takes over: calibrated once per lmc_chains_begin from the
    // measured ray counts of one steady-state iteration (0 = not calibrated yet).  Pure scheduling: results do
    // not depend on it.
    int *fullWavesTuned;
    long long *iterationsSinceBegin;
    // Small-step and large-step chains are disjoint, so their start / post kernels (DRAM-latency bound, < 2 %
    // issue utilisation each) run side by side: the large-step one goes to this auxiliary stream.
    cudaStream_t aux;
    cudaEvent_t evFork, evJoin;
    int *cacheBlockCounts;       // [LMC_CACHE_SLOTS][ceil(n / LMC_CACHE_BLOCK)] + 1 (active flag); global cache runs only
};
#define LMC_DECLARE_CHAIN(MAXD) \
    size_t chain_state_bytes_##MAXD(); \
    size_t gen_work_bytes_##MAXD(); \
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, void *states, int n, int chainBase, const float *initLs); \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, void *states, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, const WaveCfg &wc, unsigned long long *launches, H2mcSide *sides); \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const void *states, int n, unsigned long long *out); \
    cudaError_t launch_mlt_init_paths_##MAXD(cudaStream_t st, const Scene &sc, long long numInitSamples, int logicalThreads, \
                                             int tBegin, int tEnd, int emit, int *counts, const long long *offsets, float *scores);
LMC_DECLARE_CHAIN(4)
LMC_DECLARE_CHAIN(8)
LMC_DECLARE_CHAIN(12)

// LMC_PHASE_TIMING=1: CUDA-event timing of the phase groups of the LAST iteration of a launch, printed to
// stderr (development aid; events are only recorded when the variable is set).
struct PhaseTimer {
    enum { MAXP = 12 };
    cudaEvent_t ev[MAXP]; const char *name[MAXP]; int n; bool on; cudaStream_t st;
    PhaseTimer(cudaStream_t s) : n(0), on(false), st(s) { const char *v = getenv("LMC_PHASE_TIMING"); on = v && v[0] == '1'; }
    void arm(bool last) { if (on && last) n = 0; else if (on) n = -1; }
    void mark(const char *what) {
        if (!on || n < 0 || n >= MAXP) return;
        cudaEventCreate(&ev[n]); cudaEventRecord(ev[n], st); name[n] = what; n++;
    }
    void report() {
        if (!on || n <= 1) return;
        cudaEventSynchronize(ev[n - 1]);
        float tot = 0.0f;
        for (int i = 1; i < n; i++) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); tot += ms; fprintf(stderr, "[lmc phase] %-22s %8.3f ms\n", name[i], ms); }
        fprintf(stderr, "[lmc phase] %-22s %8.3f ms\n", "iteration", tot);
        for (int i = 0; i < n; i++) cudaEventDestroy(ev[i]);
    }
};

// H2MC: Hessian kernel (k_wave_grad<MAXD, 2>) + cooperative Gaussian kernel over one gradient list.  Defined in
// chain_hess_<MAXD>.cu (LMC_INSTANTIATE_HESS): the dual-number reverse sweep is the longest ptxas job of the library, so it
// gets translation units of its own that build next to chain_inst_<MAXD>.cu.
template <int MAXD>
cudaError_t launch_wave_hess(cudaStream_t st, const Scene &sc, ChainRec<MAXD> *states, int n, const int *list, const int *count,
                             int which, H2mcSide *sides, H2mcSide *padSide, int GG, int GH);
template <> cudaError_t launch_wave_hess<4>(cudaStream_t, const Scene &, ChainRec<4> *, int, const int *, const int *, int, H2mcSide *, H2mcSide *, int, int);
template <> cudaError_t launch_wave_hess<8>(cudaStream_t, const Scene &, ChainRec<8> *, int, const int *, const int *, int, H2mcSide *, H2mcSide *, int, int);
template <> cudaError_t launch_wave_hess<12>(cudaStream_t, const Scene &, ChainRec<12> *, int, const int *, const int *, int, H2mcSide *, H2mcSide *, int, int);
#define LMC_INSTANTIATE_HESS(MAXD) \
    template <> cudaError_t launch_wave_hess<MAXD>(cudaStream_t st, const Scene &sc, ChainRec<MAXD> *states, int n, const int *list, \
                                                   const int *count, int which, H2mcSide *sides, H2mcSide *padSide, int GG, int GH) { \
        k_wave_grad<MAXD, 2><<<GG * (LMC_GRAD_BLOCK / LMC_HESS_BLOCK), LMC_HESS_BLOCK, 0, st>>>(sc, states, n, list, count, which, sides, padSide); \
        k_h2mc_gaussian<MAXD><<<GH, LMC_H2MC_BLOCK, 0, st>>>(sc, states, list, count, which, sides); \
        return cudaGetLastError(); \
    }

// One iteration of the chain loop for all chains = the launch sequence below.
template <int MAXD>
cudaError_t launch_chain_run_t(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, void *states_,
                               int n, long long numSteps, float *film, unsigned char *trace, float *aTrace,
                               const WaveLists &wl, const WaveCfg &wc, unsigned long long *launches, H2mcSide *sides) {
    typedef typename GenWorkT<MAXD>::type GW;
    WaveQueues wq = wc.wq;       // + what the POST part of the small steps needs
    wq.rp = rp; wq.propGrad = wl.propGrad;
    ChainRec<MAXD> *states = (ChainRec<MAXD> *)states_;
    GW *genWork = (GW *)wc.genWork;
    const int B = LMC_CHAIN_BLOCK, G = (n + B - 1) / B;
    const int sms = wc.smCount > 0 ? wc.smCount : 148;
    const int GSmax = (n + LMC_SHADE_BLOCK - 1) / LMC_SHADE_BLOCK;
    const int GS = GSmax < sms * LMC_SHADE_MINB ? GSmax : sms * LMC_SHADE_MINB;     // grid-stride kernels: one resident wave of CTAs
    const int GSS = GSmax < sms * LMC_START_MINB ? GSmax : sms * LMC_START_MINB;
    const int GTmax = (n + LMC_TRACE_BLOCK - 1) / LMC_TRACE_BLOCK;
    const int GT = GTmax < sms * LMC_TRACE_MINB ? GTmax : sms * LMC_TRACE_MINB;     // persistent traversal warps
    const int maxDepth = sc.opt.maxDepth;
    const bool useWavefront = wc.wavefront < 0 ? (n >= LMC_WAVEFRONT_MIN_CHAINS) : (wc.wavefront != 0);
    const int GALIGN = LMC_GRAD_BLOCK;          // gradient lists: class-pure blocks
    const int GG = (n + LMC_NKEYS * (GALIGN - 1) + LMC_GRAD_BLOCK - 1) / LMC_GRAD_BLOCK;   // gradient grid over the (padded) list
    const int GHmax = (n + LMC_NKEYS * (GALIGN - 1)) / (LMC_H2MC_BLOCK / 16) + 1;
    const int GH = GHmax < sms * 8 ? GHmax : sms * 8;                                   // H2MC Gaussians: 16 lanes per list entry, grid-stride
    cudaError_t e = cudaMemsetAsync(wl.largeCount, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    k_wave_begin<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl);
    *launches += 1;
    PhaseTimer pt(st);
    for (long long k = 0; k < numSteps; k++) {
        pt.arm(k + 1 == numSteps);
        pt.mark("start");
        // Monolithic form (small jobs): the large-step proposals touch only their own chains and depend on nothing this
        // iteration has produced yet, and a small job leaves most of the SMs idle -- they run on the auxiliary stream next to
        // the current-state gradients and the small-step proposals (2^16 chains: k_wave_propose<LARGE> 0.51 ms for 8 k
        // chains, one thread per whole path, against 0.50 ms for the 57 k small steps).
        const bool largeAside = !useWavefront && wc.aux != nullptr;
        if (largeAside) {
            cudaEventRecord(wc.evFork, st); cudaStreamWaitEvent(wc.aux, wc.evFork, 0);
            k_wave_propose<MAXD, 0><<<G, B, 0, wc.aux>>>(sc, rp, chainBase, states, n, wl.large, wl.largeCount, wl, sides);
            cudaEventRecord(wc.evJoin, wc.aux);
        }
        k_sort_scan<<<2, 1024, 0, st>>>(wl.small_, wl.curGrad, 2, 1, GALIGN);
        if (GALIGN > 1) {
            e = cudaMemsetAsync(wl.curGrad.list, 0xFF, sizeof(int) * (size_t)wl.listLen, st);
            if (e != cudaSuccess) return e;
        }
        k_sort_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, wl.small_, wl.curGrad, 2);
        if (sc.opt.h2mc) {
            e = launch_wave_hess<MAXD>(st, sc, states, n, wl.curGrad.list, wl.curGrad.count, 0, sides, wc.padSide, GG, GH);
            if (e != cudaSuccess) return e;
            *launches += 1;
        } else k_wave_grad<MAXD, 1><<<GG, LMC_GRAD_BLOCK, 0, st>>>(sc, states, n, wl.curGrad.list, wl.curGrad.count, 0, sides, wc.padSide);
        *launches += 3;
        pt.mark("sort + grad(cur)");
        if (!useWavefront) {
            k_wave_propose<MAXD, 1><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.small_.list, wl.small_.count, wl, sides);
            if (largeAside) cudaStreamWaitEvent(st, wc.evJoin, 0);
            else k_wave_propose<MAXD, 0><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl.large, wl.largeCount, wl, sides);
            *launches += 2;
        } else {
            e = cudaMemsetAsync(wc.queueCounts, 0, LMC_NCOUNTERS * sizeof(int), st);
            if (e != cudaSuccess) return e;
            cudaStream_t sa = wc.aux ? wc.aux : st;
            if (wc.aux) { cudaEventRecord(wc.evFork, st); cudaStreamWaitEvent(wc.aux, wc.evFork, 0); }
            k_prop_start<MAXD, 1><<<GSS, LMC_SHADE_BLOCK, 0, sa>>>(sc, chainBase, states, genWork, wl.large, wl.largeCount, wq, sides);
            if (wc.aux) cudaEventRecord(wc.evJoin, wc.aux);
            k_prop_start<MAXD, 0><<<GSS, LMC_SHADE_BLOCK, 0, st>>>(sc, chainBase, states, genWork, wl.small_.list, wl.small_.count, wq, sides);
            if (wc.aux) cudaStreamWaitEvent(st, wc.evJoin, 0);
            *launches += 2;
            pt.mark("prop_start");
            // a path has at most maxDepth - 1 light-subpath and maxDepth camera-subpath vertices
            const int numWaves = 2 * maxDepth - 1;
            // calibration iteration (the 4th after begin: step mix and path lengths have settled): all waves are
            // full waves and the host reads the number of rays left after each one
            const bool calibrate = *wc.fullWavesTuned == 0 && *wc.iterationsSinceBegin == 3;
            int fullWaves = *wc.fullWavesTuned > 0 ? *wc.fullWavesTuned : LMC_FULL_WAVES;
            if (calibrate || fullWaves > numWaves) fullWaves = numWaves;
            int calibrated = numWaves;
            for (int w = 0; w < fullWaves; w++) {
                const int cur = w & 1;
                k_trace<<<GT, LMC_TRACE_BLOCK, 0, st>>>(sc, wq, cur, wc.queueCounts + 16 + w);
                if (w < maxDepth - 1) {
                    k_shade<MAXD, TS_P_LGT><<<GS, LMC_SHADE_BLOCK, 0, st>>>(sc, chainBase, states, genWork, wq, cur);
                    k_shade<MAXD, TS_G_LGT><<<GS, LMC_SHADE_BLOCK, 0, st>>>(sc, chainBase, states, genWork, wq, cur);
                    *launches += 2;
                }
                k_shade<MAXD, TS_P_CAM><<<GS, LMC_SHADE_BLOCK, 0, st>>>(sc, chainBase, states, genWork, wq, cur);
                k_shade<MAXD, TS_G_CAM><<<GS, LMC_SHADE_BLOCK, 0, st>>>(sc, chainBase, states, genWork, wq, cur);
                *launches += 3;
                e = cudaMemsetAsync(wc.queueCounts + 4 * cur, 0, 4 * sizeof(int), st);
                if (e != cudaSuccess) return e;
                if (calibrate && calibrated == numWaves) {
                    int left[4];
                    e = cudaMemcpyAsync(left, wc.queueCounts + 4 * (cur ^ 1), sizeof(left), cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                    if (e != cudaSuccess) return e;
                    // hand over to the tail kernel once less than 1/LMC_TAIL_DIV of the chains still have a ray in flight
                    if ((long long)left[0] + left[1] + left[2] + left[3] < (long long)n / LMC_TAIL_DIV && w + 1 >= 3) calibrated = w + 1;
                }
            }
            if (calibrate) {
                *wc.fullWavesTuned = calibrated;
                if (pt.on) fprintf(stderr, "[lmc phase] calibrated: %d full waves of %d, then the tail kernel\n", calibrated, numWaves);
            }
            pt.mark("waves (trace + shade)");
            if (fullWaves < numWaves) {
                k_shade_tail<MAXD><<<G, 128, 0, st>>>(sc, chainBase, states, genWork, wq, fullWaves & 1);
                *launches += 1;
            }
            pt.mark("tail");
            k_connect<MAXD><<<G, 128, 0, st>>>(sc, states, genWork, wq);
            *launches += 1;
            pt.mark("connect");
            k_shadow<MAXD><<<GT, LMC_TRACE_BLOCK, 0, st>>>(sc, states, wq, wc.queueCounts + 16 + 63);
            // (the POST part of the small steps ran where their candidate was resolved: k_shade / k_shadow)
            k_prop_post_large<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, genWork, wl.large, wl.largeCount, wl);
            *launches += 2;
            pt.mark("shadow + post");
        }
        k_sort_scan<<<1, 1024, 0, st>>>(wl.propGrad, wl.propGrad, 1, GALIGN, GALIGN);
        if (GALIGN > 1) {
            e = cudaMemsetAsync(wl.propGrad.list, 0xFF, sizeof(int) * (size_t)wl.listLen, st);
            if (e != cudaSuccess) return e;
        }
        k_sort_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, wl.propGrad, wl.propGrad, 1);
        if (sc.opt.h2mc) {
            e = launch_wave_hess<MAXD>(st, sc, states, n, wl.propGrad.list, wl.propGrad.count, 1, sides, wc.padSide, GG, GH);
            if (e != cudaSuccess) return e;
            *launches += 1;
        } else k_wave_grad<MAXD, 1><<<GG, LMC_GRAD_BLOCK, 0, st>>>(sc, states, n, wl.propGrad.list, wl.propGrad.count, 1, sides, wc.padSide);
        pt.mark("sort + grad(prop)");
        // the large-step list of this iteration is consumed: refill it for the next one in the fused finish + begin
        e = cudaMemsetAsync(wl.largeCount, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
        if (sc.opt.cacheEnabled) {
            // chains interact through the cache: the iteration's push requests are committed (in chain order) between
            // its finish and the next iteration's begin, so the two are not fused
            const int CB = (n + LMC_CACHE_BLOCK - 1) / LMC_CACHE_BLOCK;
            int *active = wc.cacheBlockCounts + LMC_CACHE_SLOTS * CB;
            k_cache_prequery<MAXD><<<(n + 255) / 256, 256, 0, st>>>(sc, states, n);
            k_wave_finish<MAXD, 0><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, film, trace, aTrace, numSteps, k, sides, wl);
            k_cache_count<MAXD><<<CB, LMC_CACHE_BLOCK, 0, st>>>(sc, states, n, wc.cacheBlockCounts);
            k_cache_scan<<<1, 1024, 0, st>>>(sc.gc, wc.cacheBlockCounts, CB, active);
            k_cache_write<MAXD><<<CB, LMC_CACHE_BLOCK, 0, st>>>(sc, states, n, wc.cacheBlockCounts, active);
            k_cache_grid<<<LMC_CACHE_SLOTS, 256, 0, st>>>(sc.gc);
            *launches += 2;
            if (k + 1 < numSteps) k_wave_begin<MAXD><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, wl);
            *launches += 4;
        } else if (k + 1 < numSteps) k_wave_finish<MAXD, 1><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, film, trace, aTrace, numSteps, k, sides, wl);
        else k_wave_finish<MAXD, 0><<<G, B, 0, st>>>(sc, rp, chainBase, states, n, film, trace, aTrace, numSteps, k, sides, wl);
        *launches += 4;
        pt.mark("finish (+ begin)");
        *wc.iterationsSinceBegin += 1;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    pt.report();
    return cudaSuccess;
}

#define LMC_INSTANTIATE_CHAIN(MAXD) \
    size_t chain_state_bytes_##MAXD() { return sizeof(ChainRec<MAXD>); } \
    size_t gen_work_bytes_##MAXD() { return sizeof(GenWorkT<MAXD>::type); } \
    cudaError_t launch_chain_init_##MAXD(cudaStream_t st, void *states, int n, int chainBase, const float *initLs) { \
        k_chain_init<MAXD><<<(n + 127) / 128, 128, 0, st>>>((ChainRec<MAXD> *)states, n, chainBase, initLs); \
        return cudaGetLastError(); \
    } \
    cudaError_t launch_chain_run_##MAXD(cudaStream_t st, const Scene &sc, const RunParams &rp, int chainBase, void *states_, \
                                        int n, long long numSteps, float *film, unsigned char *trace, float *aTrace, \
                                        const WaveLists &wl, const WaveCfg &wc, unsigned long long *launches, H2mcSide *sides) { \
        return launch_chain_run_t<MAXD>(st, sc, rp, chainBase, states_, n, numSteps, film, trace, aTrace, wl, wc, launches, sides); \
    } \
    cudaError_t launch_chain_stats_##MAXD(cudaStream_t st, const void *states, int n, unsigned long long *out) { \
        k_chain_stats<MAXD><<<(n + 127) / 128, 128, 0, st>>>((const ChainRec<MAXD> *)states, n, out); \
        return cudaGetLastError(); \
    } \
    cudaError_t launch_mlt_init_paths_##MAXD(cudaStream_t st, const Scene &sc, long long numInitSamples, int logicalThreads, \
                                             int tBegin, int tEnd, int emit, int *counts, const long long *offsets, float *scores) { \
        const int g = (tEnd - tBegin + 127) / 128; \
        if (emit) k_mlt_init_paths<MAXD, 1><<<g, 128, 0, st>>>(sc, numInitSamples, logicalThreads, tBegin, tEnd, counts, offsets, scores); \
        else k_mlt_init_paths<MAXD, 0><<<g, 128, 0, st>>>(sc, numInitSamples, logicalThreads, tBegin, tEnd, counts, offsets, scores); \
        return cudaGetLastError(); \
    }

}  // namespace lmc_cuda
