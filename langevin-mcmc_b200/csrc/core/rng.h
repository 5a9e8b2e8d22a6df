// rng.h -- bit-exact re-statement of the reference's RNG and of the libstdc++ distributions
// it draws through.
//
//   RNG = pcg32_k64_fast = extended<6, 32, oneseq_xsh_rs_64_32, oneseq_rxs_m_xs_32_32, kdd=true>
//     (reference: src/commondef.h:63, src/pcg_random.hpp:1692, :1155-1217 (extended),
//      :787-809 (xsh_rs output), :425-437 (seeding), :1337-1352 (selfinit),
//      :1439-1448 (advance_table), :1114-1129 (inside_out::external_step),
//      :920-951 (rxs_m_xs output/unoutput), src/pcg_extras.hpp:255-272 (unxorshift)).
//   std::uniform_real_distribution<float>, std::normal_distribution<float> as implemented by
//     libstdc++ (bits/random.tcc: generate_canonical :3345-3381, normal_distribution :1811-1844).
//
// The 64-entry extension table lives behind a (pointer, stride) pair so the device can keep
// it in shared memory with one column per thread (bank-conflict free) while the host keeps a
// plain array.  The table is a pure function of (seed, number of table advances), so the
// persistent chain state only stores the 64-bit LCG state plus an advance counter.
#pragma once
#include "detmath.h"

namespace lmc {

#define LMC_PCG_MULT 6364136223846793005ULL
#define LMC_PCG_INC 1442695040888963407ULL

struct Rng {
    uint64_t state;
    uint32_t *tab;    // 64 entries, element i at tab[i * stride]
    int stride;
    uint32_t epoch;   // number of advance_table() calls so far (persisted with the chain)
    // Lazy mode: while the table has never advanced (epoch == 0; an advance happens once per 2^32
    // draws) entry i is a pure function of the seed, xsh_rs(s_{i+2}) ^ xdiff, and is computed on
    // demand with the LCG jump-ahead constants below instead of materialising 64 entries.
    int lazy;
    uint64_t s0;      // bump(seed + increment): the state selfinit() starts from
    uint32_t xdiff;
};

// (A^k, C_k), k = 0..66: s_k = A^k * s_0 + C_k
#define LMC_PCG_JUMP_N 67
static const unsigned long long LMC_PCG_JUMP_H[2 * LMC_PCG_JUMP_N] = {
#include "pcg_jump.inc"
};
#if defined(__CUDACC__)
static __device__ const unsigned long long LMC_PCG_JUMP_D[2 * LMC_PCG_JUMP_N] = {
#include "pcg_jump.inc"
};
#endif
LMC_HD uint64_t pcg_jump(uint64_t s0, uint32_t k) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 ac = __ldg(reinterpret_cast<const ulonglong2 *>(LMC_PCG_JUMP_D) + k);
    return ac.x * s0 + ac.y;
#else
    return LMC_PCG_JUMP_H[2 * k] * s0 + LMC_PCG_JUMP_H[2 * k + 1];
#endif
}

LMC_HD uint32_t pcg_xsh_rs(uint64_t s) {
    const uint32_t rshift = (uint32_t)(s >> 61) & 7u;
    s ^= s >> 22;
    return (uint32_t)(s >> (22u + rshift));
}

LMC_HD uint32_t pcg_base_next(uint64_t &state) {
    const uint64_t old = state;
    state = old * LMC_PCG_MULT + LMC_PCG_INC;
    return pcg_xsh_rs(old);
}

// --- rxs_m_xs 32/32 permutation and its inverse (used only by advance_table) ---
LMC_HD uint32_t pcg_rxs_m_xs32_output(uint32_t v) {
    const uint32_t rshift = (v >> 28) & 15u;
    v ^= v >> (4u + rshift);
    v *= 277803737u;
    v ^= v >> 22;
    return v;
}
LMC_HD uint32_t pcg_unxorshift32(uint32_t x, uint32_t bits, uint32_t shift) {
    // iterative form of the recursive inverse: x = y ^ (y >> shift)
    uint32_t y = x;
    for (uint32_t s = shift; s < bits; s += shift) {
        y = x ^ (y >> shift);
    }
    return y;
}
LMC_HD uint32_t pcg_rxs_m_xs32_unoutput(uint32_t v) {
    v = pcg_unxorshift32(v, 32, 22);
    v *= 2897767785u;
    const uint32_t rshift = (v >> 28) & 15u;
    v = pcg_unxorshift32(v, 32, 4u + rshift);
    return v;
}
LMC_HD bool pcg_external_step(uint32_t &randval, uint32_t i) {
    uint32_t st = pcg_rxs_m_xs32_unoutput(randval);
    st = st * 747796405u + 2891336453u + i * 2u;
    const uint32_t result = pcg_rxs_m_xs32_output(st);
    randval = result;
    return result == 0u;
}

LMC_HD void rng_materialize(Rng &r) {      // leave lazy mode: write the 64 seed-defined entries
    if (!r.lazy) return;
    for (uint32_t i = 0; i < 64u; ++i) r.tab[i * r.stride] = pcg_xsh_rs(pcg_jump(r.s0, i + 2u)) ^ r.xdiff;
    r.lazy = 0;
}

LMC_HD void rng_advance_table(Rng &r) {
    rng_materialize(r);
    bool carry = false;
    for (uint32_t i = 0; i < 64u; ++i) {
        uint32_t v = r.tab[i * r.stride];
        if (carry) {
            carry = pcg_external_step(v, i + 1u);
        }
        const bool carry2 = pcg_external_step(v, i + 1u);
        carry = carry || carry2;
        r.tab[i * r.stride] = v;
    }
    r.epoch += 1u;
}

// RNG(seed): state = bump(seed + increment); then selfinit() fills the table from the base
// generator.  `xdiff = base() - base()`: gcc evaluates the left operand first.
LMC_HD void rng_seed(Rng &r, uint64_t seed) {
    r.state = (seed + LMC_PCG_INC) * LMC_PCG_MULT + LMC_PCG_INC;
    const uint32_t a = pcg_base_next(r.state);
    const uint32_t b = pcg_base_next(r.state);
    const uint32_t xdiff = a - b;
    for (int i = 0; i < 64; ++i) {
        r.tab[i * r.stride] = pcg_base_next(r.state) ^ xdiff;
    }
    r.epoch = 0u;
    r.lazy = 0;
}

// Lazy forms of RNG(seed) and of the restore below: nothing is written to the table unless it
// has to advance (r.tab must still point at 64 words of scratch for that case).
LMC_HD void rng_lazy_base(Rng &r, uint64_t seed) {
    r.s0 = (seed + LMC_PCG_INC) * LMC_PCG_MULT + LMC_PCG_INC;
    r.xdiff = pcg_xsh_rs(r.s0) - pcg_xsh_rs(pcg_jump(r.s0, 1u));
    r.lazy = 1;
}
LMC_HD void rng_seed_lazy(Rng &r, uint64_t seed) {
    rng_lazy_base(r, seed);
    r.state = pcg_jump(r.s0, 66u);
    r.epoch = 0u;
}
LMC_HD void rng_advance_table(Rng &r);
LMC_HD void rng_restore_lazy(Rng &r, uint64_t seed, uint64_t state, uint32_t epoch) {
    rng_lazy_base(r, seed);
    r.epoch = 0u;
    for (uint32_t e = 0; e < epoch; ++e) rng_advance_table(r);
    r.state = state;
}

// Rebuild the table for a chain that was seeded with `seed` and has advanced its table
// `epoch` times; `state` is the persisted LCG state.
LMC_HD void rng_restore(Rng &r, uint64_t seed, uint64_t state, uint32_t epoch) {
    rng_seed(r, seed);
    for (uint32_t e = 0; e < epoch; ++e) rng_advance_table(r);
    r.epoch = epoch;
    r.state = state;
}

LMC_HD uint32_t rng_next(Rng &r) {
    const uint64_t s = r.state;
    if ((s & 0xFFFFFFFFULL) == 0ULL) {
        rng_advance_table(r);
    }
    const uint32_t idx = (uint32_t)(s & 63ULL);
    const uint32_t rhs = r.lazy ? (pcg_xsh_rs(pcg_jump(r.s0, idx + 2u)) ^ r.xdiff) : r.tab[idx * r.stride];
    const uint32_t lhs = pcg_base_next(r.state);
    return lhs ^ rhs;
}

// std::generate_canonical<float, 24>(rng): one 32-bit draw.
LMC_HD float rng_canonical(Rng &r) {
    const float sum = (float)rng_next(r);
    float ret = sum / 4294967296.0f;
    if (ret >= 1.0f) ret = 0.99999994f;  // nextafterf(1, 0)
    return ret;
}

// std::uniform_real_distribution<float>(a, b)(rng)
LMC_HD float rng_uniform(Rng &r) { return rng_canonical(r) * (1.0f - 0.0f) + 0.0f; }
LMC_HD float rng_uniform_ab(Rng &r, float a, float b) { return rng_canonical(r) * (b - a) + a; }

// std::normal_distribution<float>: Marsaglia polar, second variate cached per object.
struct NormalDist {
    float mean, stddev;
    float saved;
    bool savedAvailable;
};
LMC_HD NormalDist normal_make(float mean, float stddev) {
    NormalDist d; d.mean = mean; d.stddev = stddev; d.saved = 0.0f; d.savedAvailable = false;
    return d;
}
LMC_HD float normal_draw(NormalDist &d, Rng &r) {
    float ret;
    if (d.savedAvailable) {
        d.savedAvailable = false;
        ret = d.saved;
    } else {
        float x, y, r2;
        do {
            x = 2.0f * rng_canonical(r) - 1.0f;
            y = 2.0f * rng_canonical(r) - 1.0f;
            r2 = x * x + y * y;
        } while (r2 > 1.0f || r2 == 0.0f);
        const float mult = dm_sqrt(-2.0f * dm_log(r2) / r2);
        d.saved = x * mult;
        d.savedAvailable = true;
        ret = y * mult;
    }
    return ret * d.stddev + d.mean;
}

}  // namespace lmc
