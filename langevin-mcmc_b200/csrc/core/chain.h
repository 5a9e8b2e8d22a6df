// chain.h -- per-chain persistent state and the K-iteration chain runner shared by the sm_100a
// kernel (one thread = one chain) and the host twin used as the bit-level oracle.
//
// Reference: the body of the ParallelFor lambda, src/mlt.cpp:60-196 (RNG(chainId + seedOffset)
// :61-62; currentState = initStates[chainId], valid = false :68, src/mlt.h:124; loop :91-170).
#pragma once
#include "mutation.h"
#include "direct.h"

namespace lmc {

template <int MAXD>
struct ChainState {
    MarkovState<MAXD> st[2];
    ChainVars<MAXD> ch;
    StepScratch<MAXD> ss;
    PropCand pc;                 // small step: the proposal's candidate contribution (wavefront, stages.h)
    LpsFull lpsFull;             // small step with a light subpath: its final state, parked for ConnectVertex
    int curIdx;
    unsigned long long rngState;
    unsigned int rngEpoch;
    unsigned int seeded;
    long long sampleIdx;
    // counters (lmc_stats)
    unsigned int nAccept[4];     // per MutationType
    unsigned int nPropose[4];
    unsigned int gradStats[2];   // gradient evaluations, non-finite gradients
};

template <int MAXD>
LMC_HD void chain_state_init(ChainState<MAXD> &cs, float initLsScore) {
    for (int k = 0; k < 2; k++) {
        MarkovState<MAXD> &s = cs.st[k];
        s.valid = 0; s.gaussianInitialized = 0; s.nSplat = 0; s.scoreSum = 0.0f;
        s.sp.camDepth = 0; s.sp.lightDepth = 0; s.sp.screenPos = mk2(0, 0); s.sp.contrib = mk3s(0.0f);
        s.sp.lsScore = 0.0f; s.sp.ssScore = 0.0f;
        path_clear(s.path);
        s.path.camDepth = 0; s.path.lgtDepth = 0; s.path.time = 0.0f;
        s.gaussian.dim = 0; s.gaussian.logDet = 0.0f;
    }
    cs.st[0].sp.lsScore = initLsScore;   // initStates[chainId].spContrib.lsScore
    chain_vars_init(cs.ch);
    cs.ss.kind = STEP_LARGE; cs.ss.needCurGrad = 0; cs.ss.needPropGrad = 0; cs.ss.hasContrib = 0; cs.ss.a = 0.0f;
    for (int i = 0; i < Limits<MAXD>::DIM; i++) { cs.ss.offset[i] = 0.0f; cs.ss.grad[i] = 0.0f; }
    cs.pc.n = 0;
    cs.curIdx = 0; cs.rngState = 0ULL; cs.rngEpoch = 0u; cs.seeded = 0u; cs.sampleIdx = 0;
    for (int i = 0; i < 4; i++) { cs.nAccept[i] = 0; cs.nPropose[i] = 0; }
    cs.gradStats[0] = 0; cs.gradStats[1] = 0;
}

// Runs `numSteps` iterations of chain `globalChainId`.  `tab`/`stride` is scratch for the
// 64-entry PCG extension table (regenerated from the seed, never stored with the chain).
// trace (optional): one byte per step = mutationType | accepted << 2 | (a > 0) << 3;
// aTrace (optional): the acceptance probability of each step.
template <int MAXD, class FILM>
LMC_HD void chain_run(const Scene &sc, const RunParams &rp, int globalChainId, ChainState<MAXD> &cs,
                      long long numSteps, uint32_t *tab, int stride, FILM &film,
                      unsigned char *trace, float *aTrace, long long traceStride, H2mcSide *side = nullptr,
                      StagedWork<MAXD> *staged = nullptr) {
    Rng rng; rng.tab = tab; rng.stride = stride;
    const uint64_t seed = (uint64_t)(long long)(globalChainId + sc.opt.seedOffset);
    if (!cs.seeded) {
        rng_seed(rng, seed);
        cs.seeded = 1u;
    } else {
        rng_restore(rng, seed, cs.rngState, cs.rngEpoch);
    }
    for (long long k = 0; k < numSteps; k++) {
        const StepInfo info = chain_step(sc, rp, globalChainId, cs.sampleIdx, cs.st, cs.curIdx, cs.ch, rng, film,
                                         cs.gradStats, cs.ss, side, staged);
        cs.nPropose[info.mutationType] += 1u;
        cs.nAccept[info.mutationType] += (unsigned int)info.accepted;
        if (trace) trace[k * traceStride] = (unsigned char)(info.mutationType | (info.accepted << 2) | ((info.a > 0.0f) ? 8 : 0));
        if (aTrace) aTrace[k * traceStride] = info.a;
        cs.sampleIdx += 1;
    }
    cs.rngState = rng.state;
    cs.rngEpoch = rng.epoch;
}

// Bin the entries of a ready slot into the query grid (scene.h); the device form is k_cache_grid.
inline void cache_grid_build_host(const GlobalCacheView &gc, int s) {
    const int dim = 4 + 2 * s;
    int *cellStart = gc.grid + (size_t)s * LMC_CACHE_GRID_INTS, *cursor = cellStart + LMC_CACHE_CELLS + 1, *entry = cursor + LMC_CACHE_CELLS;
    const float *base = gc.data + cache_slot_offset(s);
    for (int c = 0; c <= LMC_CACHE_CELLS; c++) cellStart[c] = 0;
    for (int e = 0; e < LMC_CACHE_MAX_SIZE; e++) cellStart[cache_cell(base + (size_t)e * 3 * dim) + 1]++;
    for (int c = 0; c < LMC_CACHE_CELLS; c++) { cellStart[c + 1] += cellStart[c]; cursor[c] = 0; }
    for (int e = 0; e < LMC_CACHE_MAX_SIZE; e++) { const int c = cache_cell(base + (size_t)e * 3 * dim); entry[cellStart[c] + cursor[c]++] = e; }
    gc.gridReady[s] = 1;
}

// Global cache: apply the push requests of ONE iteration in chain order (src/mlt.cpp:121-127 pushes under a mutex in
// whatever order the threads arrive; here the order is defined: by chain id, all requests of an iteration after the
// iteration).  Entries beyond PSS_MAX_SIZE are dropped (global_cache_t::push returns false once is_ready), a slot becomes
// ready when it holds PSS_MAX_SIZE entries.  Host form; the device runs the same rule as k_cache_count / _scan / _write.
template <int MAXD>
inline void cache_commit_host(const Scene &sc, ChainState<MAXD> *const *states, int n) {
    for (int i = 0; i < n; i++) {
        ChainVars<MAXD> &ch = states[i]->ch;
        const int dim = ch.pushDim;
        if (dim <= 0) continue;
        ch.pushDim = 0;
        const int s = cache_slot(dim);
        if (s < 0 || sc.gc.count[s] >= LMC_CACHE_MAX_SIZE) continue;
        float *e = sc.gc.data + cache_slot_offset(s) + (size_t)sc.gc.count[s] * 3 * dim;
        for (int k = 0; k < dim; k++) { e[k] = ch.pss[k]; e[dim + k] = ch.v1[k]; e[2 * dim + k] = ch.v2[k]; }
        sc.gc.count[s] += 1;
    }
    for (int s = 0; s < LMC_CACHE_SLOTS; s++) {
        if (sc.gc.count[s] >= LMC_CACHE_MAX_SIZE) sc.gc.ready[s] = 1;
        if (sc.gc.ready[s] && sc.gc.grid && !sc.gc.gridReady[s]) cache_grid_build_host(sc.gc, s);
    }
}

}  // namespace lmc
