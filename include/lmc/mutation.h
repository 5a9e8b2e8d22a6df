// mutation.h -- mirror of the reference's mutation plugin surface (src/mutation.h:5-43).
//
// In the reference a chain owns `std::unique_ptr<Mutation>` objects (LargeStep, SmallStep, MALASmallStep,
// H2MCSmallStep; src/mlt.cpp:71-90) and calls the virtual `Mutate(mltState, normalization, currentState,
// proposalState, rng, chain)` once per iteration.  On the device the mutation kinds are the values of
// MutationType, chosen per chain and iteration exactly like src/mlt.cpp:96-101 / src/mutation_mala.h:47-51, and
// `Mutate` is the kernel sequence of one lmc_run_chains iteration (csrc/cuda/chain_kernels.cuh):
//   Large      GeneratePathBidir + technique selection          src/mutation_large.h:31-127
//   Small      isotropic Gaussian in primary sample space       src/mutation_small.h:16-55
//   MALASmall  Langevin proposal, diagonal preconditioner       src/mutation_mala.h:35-278, src/mala.cpp:7-51
//   H2MCSmall  Hessian-Hamiltonian proposal, dense Gaussian     src/mutation_h2mc.h:38-127, src/h2mc.cpp:3-142
// Which small step runs is selected by DptOptions::mala / ::h2mc like src/mlt.cpp:71-85.
#pragma once
#include <stdint.h>
#include <vector>
#include "dptoptions.h"
#include "lmc_abi.h"

namespace lmc {

// src/mutation.h:5-8 (run-time options here: DptOptions::outlier*)
#define LMC_OUTLIER_WEAK_REJECT_CNT 10000
#define LMC_OUTLIER_STRONG_REJECT_CNT 1000
#define LMC_OUTLIER_RATIO_THRESHOLD 30.0f

enum class MutationType { Large, Small, H2MCSmall, MALASmall };    // src/mutation.h:11; index of lmc_stats.proposed[]

// One entry of the optional decision trace of lmc_run_chains: what a chain did in one iteration.
struct MutationRecord {
    MutationType lastMutationType;   // Mutation::lastMutationType, src/mutation.h:25
    bool accepted;                   // u <= a, src/mlt.cpp:113
    bool positive;                   // a > 0 (a proposal with a contribution)
    static MutationRecord Decode(uint8_t b) {
        MutationRecord r;
        r.lastMutationType = (MutationType)(b & 3); r.accepted = ((b >> 2) & 1) != 0; r.positive = ((b >> 3) & 1) != 0;
        return r;
    }
};

// Per-chain adaptation state of the reference (`struct Chain`, src/mutation.h:28-43): the Adam-style moments v1 / v2 of
// the PSS gradient, the step counter t and `buffered`.  It lives in HBM inside the chain record (csrc/core/mutation.h
// ChainVars) and never crosses the ABI; this struct documents the correspondence for readers of the reference:
//   Chain::v1, v2, curr_new_*, prop_new_*   ChainVars::v1, v2, curr_new_v1/2, prop_new_v1/2   (2 * maxDepth floats each)
//   Chain::buffered, t                      ChainVars::buffered, t
//   Chain::pss, last_pss, queried           ChainVars::pss, last_pss, queried: the global cache (DptOptions::globalCache)
struct Chain {
    int chainId = 0;
    int t = 0;
    bool buffered = false;
};

}  // namespace lmc
