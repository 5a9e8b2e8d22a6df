// mutation.h -- Gaussian proposals, the three mutation kernels and one iteration of the MLT
// chain loop.  The reference's virtual `Mutation::Mutate` (src/mutation.h:16-26) becomes the
// enum-dispatched `chain_step`; one call == one iteration of the loop at src/mlt.cpp:91-170.
//
//   Gaussian / IsotropicGaussian / GaussianLogPdf / GenerateSample   src/gaussian.{h,cpp}
//   ComputeGaussian (LMC, diagonal)                                  src/mala.cpp:7-51
//   SmallStep::Mutate                                                src/mutation_small.h:16-55
//   MALASmallStep::Mutate                                            src/mutation_mala.h:35-278
//   LargeStep::Mutate                                                src/mutation_large.h:31-127
//   chain loop body (step choice, splat, accept, moment commit, outlier reset)  src/mlt.cpp:91-170
//   Splat                                                            src/image.h:66-77
//
// The global KD-tree cache is out of scope: `isReady(dim)` is constant
// false, so every eligible MALA step evaluates a gradient.
#pragma once
#include "path.h"
#include "stages.h"
#include "serialize.h"

namespace lmc {

enum MutationType { MUT_LARGE = 0, MUT_SMALL = 1, MUT_H2MC_SMALL = 2, MUT_MALA_SMALL = 3 };  // src/mutation.h:11

#define LMC_PCD_MIN 0.01f
#define LMC_PCD_MAX 100.0f
#define LMC_MTM_MIN (-5.0f)
#define LMC_MTM_MAX 5.0f
// outlier thresholds (src/mutation.h:5-8): Options::outlier* (defaults 10000 / 1000 / 30)

template <int DIM>
struct Gaussian {                // src/gaussian.h:9-19 (diagonal members only; dense = H2MC)
    int dim;
    float logDet;
    float mean[DIM], covL_d[DIM], invCov_d[DIM];
};

template <int DIM>
LMC_HD void isotropic_gaussian(int dim, float sigma, Gaussian<DIM> &g) {
    g.dim = dim;
    const float inv = 1.0f / (sigma * sigma);
    for (int i = 0; i < dim; i++) { g.mean[i] = 0.0f; g.covL_d[i] = sigma; g.invCov_d[i] = inv; }
    g.logDet = (float)dim * dm_fastlog(inverse(sigma * sigma));
}

template <int DIM>
LMC_HD float gaussian_log_pdf(const float *offset, float sign, const Gaussian<DIM> &g) {
    const int dim = g.dim;
    float logPdf = (float)dim * (-0.9189385332046727f);
    logPdf += 0.5f * g.logDet;
    float q = 0.0f;
    for (int i = 0; i < dim; i++) {
        const float d = sign * offset[i] - g.mean[i];
        q += d * (g.invCov_d[i] * d);
    }
    logPdf -= 0.5f * q;
    return logPdf;
}

template <int DIM>
LMC_HD void generate_sample(const Gaussian<DIM> &g, float *x, Rng &rng) {
    // (g and x both live in the chain record: the dimension is read once and each element is scaled where it is drawn, so
    // that the loop neither reloads its bound after every store nor makes a second pass through global memory -- same
    // draws, same arithmetic as `x = normal; x = covL_d * x + mean`, src/gaussian.cpp:44-54)
    NormalDist nd = normal_make(0.0f, 1.0f);
    const int dim = g.dim;
    for (int i = 0; i < dim; i++) {
        const float cl = g.covL_d[i], mu = g.mean[i];       // in flight while the variate is drawn
        const float z = normal_draw(nd, rng);
        x[i] = cl * z + mu;
    }
}

// ComputeGaussian (src/mala.cpp:7-51)
template <int DIM>
LMC_HD void compute_gaussian_lmc(int dim, const float *v1, const float *M, float ss, float shk, float sc,
                                 Gaussian<DIM> &g) {
    g.dim = dim;
    const float shrk = inverse(shk * shk);
    if (sc <= 1e-10f) {
        for (int i = 0; i < dim; i++) { g.mean[i] = 0.0f; g.invCov_d[i] = shrk; g.covL_d[i] = shk; }
        g.logDet = (float)dim * dm_fastlog(inverse(shk * shk));
    } else {
        float logDet = 0.0f;            // summed in a register (g is the chain record), stored once
        for (int i = 0; i < dim; i++) {
            const float cov_t = ss * ss * (M[i] + 1.0f);
            const float invcov = inverse(cov_t) + shrk;
            const float cov = inverse(invcov);
            g.invCov_d[i] = invcov;
            g.covL_d[i] = dm_sqrt(cov);
            g.mean[i] = dm_clamp(v1[i], LMC_MTM_MIN, LMC_MTM_MAX) * cov / 2.0f;
            logDet += dm_fastlog(invcov);
        }
        g.logDet = logDet;
    }
}

struct SplatSample { V2 screenPos; V3 contrib; };

template <int MAXD>
struct Limits {
    static const int DIM = 2 * MAXD;
    // upper bound on the contributions of one GeneratePathBidir call (path.h derivation)
    static const int MAXC = (MAXD - 1) + 1 + (MAXD - 1) + ((MAXD - 2) * (MAXD - 1)) / 2 + 2;
};

template <int MAXD>
struct MarkovState {             // src/mlt.h:30-39
    int valid;
    SubpathContrib sp;
    Path<MAXD> path;
    float scoreSum;
    int gaussianInitialized;
    Gaussian<Limits<MAXD>::DIM> gaussian;
    int nSplat;
    SplatSample splat[Limits<MAXD>::MAXC];
};

template <int MAXD>
struct ChainVars {               // Chain (src/mutation.h:28-43) + per-chain loop locals (src/mlt.cpp:61-90)
    float v1[Limits<MAXD>::DIM], v2[Limits<MAXD>::DIM];
    float curr_new_v1[Limits<MAXD>::DIM], curr_new_v2[Limits<MAXD>::DIM];
    float prop_new_v1[Limits<MAXD>::DIM], prop_new_v2[Limits<MAXD>::DIM];
    int buffered;
    int t;
    float lastScoreSum, lastScore;    // LargeStep members (src/mutation_large.h:16-17)
    int adjacentReject;
    int lastMutationType;
    int outlierResets;                // statistics only: how often the reset of src/mlt.cpp:147-169 fired
    // global cache (src/mutation.h:28-43 Chain::pss, last_pss, pathWeight, queried; used only with opt.cacheEnabled)
    float pss[Limits<MAXD>::DIM], last_pss[Limits<MAXD>::DIM];
    float pathWeight;
    int queried;
    int cacheQueries, cacheHits;      // statistics: global_cache_t::query calls / calls that found a neighbour
    int pushDim;                      // > 0: this chain asks for a push of (pss, v1, v2) into the cache of that dimension
    int preq;                         // device only: the proposal's cache query of this iteration has been answered by the
                                      // warp-cooperative k_cache_prequery (1 = no neighbour, 2 = v1 / v2 already averaged in)
};

template <int MAXD>
LMC_HD void chain_vars_init(ChainVars<MAXD> &c) {
    for (int i = 0; i < Limits<MAXD>::DIM; i++) {
        c.v1[i] = 0; c.v2[i] = 0; c.curr_new_v1[i] = 0; c.curr_new_v2[i] = 0; c.prop_new_v1[i] = 0; c.prop_new_v2[i] = 0;
    }
    c.buffered = 0; c.t = 0; c.lastScoreSum = 1.0f; c.lastScore = 1.0f; c.adjacentReject = 0; c.lastMutationType = MUT_LARGE; c.outlierResets = 0;
    for (int i = 0; i < Limits<MAXD>::DIM; i++) { c.pss[i] = 0; c.last_pss[i] = 0; }
    c.pathWeight = 0.0f; c.queried = 0; c.pushDim = 0; c.cacheQueries = 0; c.cacheHits = 0; c.preq = 0;
}

// Film accumulation (src/image.h:66-77).  FILM is a functor add(pixelIndex, channel, value)
// so the device can use red.global.add.f32 and the host twin a plain +=.
template <class FILM>
LMC_HD void splat(FILM &film, int width, int height, V2 screenPos, V3 contrib) {
    const int ix = dm_clampi((int)(screenPos.x * (float)width), 0, width - 1);
    const int iy = dm_clampi((int)(screenPos.y * (float)height), 0, height - 1);
    if (all_finite(contrib)) {
        const int pix = iy * width + ix;
        film.add(pix, 0, contrib.x); film.add(pix, 1, contrib.y); film.add(pix, 2, contrib.z);
    }
}

// ------------------------------------------------------------------------------------------
// One iteration of the chain loop, split into PHASES so the device can run each phase as its
// own kernel over a compacted list of chains (wavefront execution) while the host twin calls the
// same functions back to back.  RNG draws happen in the reference's order:
//   [u_large if cur.valid] -> Mutate{ u_mix -> D normals -> PerturbPathBidir normals } -> [u_accept if a > 0]
// The gradient phases draw nothing.
// ------------------------------------------------------------------------------------------
enum StepKind { STEP_LARGE = 0, STEP_ISO = 1, STEP_MALA = 2, STEP_H2MC = 3 };

// ---- H2MC (src/mutation_h2mc.h, src/h2mc.cpp): dense Gaussians live beside the chain record -----
#define LMC_H2MC_DIM LMC_HESS_MAXDIM      // 16: derivative functions exist for c + l - 1 <= 8
struct H2mcSide {
    float hess[LMC_H2MC_DIM * LMC_H2MC_DIM];          // scratch: Hessian of the state being initialised
    float covL[2][LMC_H2MC_DIM * LMC_H2MC_DIM];       // per MarkovState slot, row-major dim x dim
    float invCov[2][LMC_H2MC_DIM * LMC_H2MC_DIM];
    int dense[2];                                      // slot holds a dense Gaussian (else the diagonal fields are used)
};

// Jacobi eigen-decomposition of a symmetric n x n matrix (n even, <= 16; row-major with row stride `ld`, destroyed).
// Replaces Eigen::SelfAdjointEigenSolver (src/h2mc.cpp:9-10; Eigen is not vendored: parity unpinned, the convention
// below is the oracle's): eigenvalues ascending, eigenvectors = columns of V, each normalised so that its first
// non-zero component is positive.
//
// The rotation order is the PARALLEL one (round-robin tournament): a sweep is n - 1 rounds, round r rotates the n / 2
// disjoint index pairs jacobi_pair(n, r, 0 .. n/2-1) at once -- all angles are taken from the matrix as it enters the
// round, then every pair's column rotation is applied, then every pair's row rotation (A <- J^T (A J), J the product
// of the round's commuting rotations).  This is the statement list both forms follow: the serial one below (host
// twin / oracle, small jobs) and the warp-cooperative one of csrc/cuda/h2mc_kernels.cuh, where 16 lanes share a
// matrix held in shared memory and the sums below are shuffle trees.  Sums over 16 lanes are defined as the xor
// butterfly tree16(): level by level x[l] += x[l ^ 8], ^ 4, ^ 2, ^ 1 (lanes beyond the data hold 0).
LMC_HD float tree16(float *x) {
    for (int off = 8; off > 0; off >>= 1) {
        float y[16];
        for (int l = 0; l < 16; l++) y[l] = x[l] + x[l ^ off];
        for (int l = 0; l < 16; l++) x[l] = y[l];
    }
    return x[0];
}
// pair i (0 <= i < n / 2) of round r (0 <= r < n - 1), p < q: circle method with n - 1 fixed
LMC_HD void jacobi_pair(int n, int r, int i, int &p, int &q) {
    const int m = n - 1;
    int a, b;
    if (i == 0) { a = m; b = r; }
    else { a = (r + i) % m; b = (r + m - i) % m; }
    p = a < b ? a : b; q = a < b ? b : a;
}
// rotation that annihilates A[p][q]; returns false when there is nothing to rotate
LMC_HD bool jacobi_angle(float app, float aqq, float apq, float &c, float &sn) {
    if (apq == 0.0f) { c = 1.0f; sn = 0.0f; return false; }
    const float theta = (aqq - app) / (2.0f * apq);
    const float t = ((theta >= 0.0f) ? 1.0f : -1.0f) / (dm_abs(theta) + dm_sqrt(theta * theta + 1.0f));
    c = 1.0f / dm_sqrt(t * t + 1.0f);
    sn = t * c;
    return true;
}
#define LMC_JACOBI_SWEEPS 16
LMC_HD bool jacobi_converged(float off, float diag) { return off <= 1e-14f * (diag + off) || off == 0.0f; }

LMC_HD_NOINLINE void jacobi_eigen(int n, float *A, float *V, float *w) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0f : 0.0f;
    const int ne = n + (n & 1);                  // an odd n plays with a phantom index whose pairs sit out (the chains only have even n)
    for (int sweep = 0; sweep < LMC_JACOBI_SWEEPS; sweep++) {
        float offp[16], diagp[16];
        for (int l = 0; l < 16; l++) {           // lane l = row l
            offp[l] = 0.0f; diagp[l] = 0.0f;
            if (l < n) { diagp[l] = A[l * n + l] * A[l * n + l]; for (int q = l + 1; q < n; q++) offp[l] += A[l * n + q] * A[l * n + q]; }
        }
        const float off = tree16(offp), diag = tree16(diagp);
        if (jacobi_converged(off, diag)) break;
        for (int r = 0; r < ne - 1; r++) {
            int P[8], Q[8]; float C[8], S[8]; bool rot[8];
            for (int i = 0; i < ne / 2; i++) {
                jacobi_pair(ne, r, i, P[i], Q[i]);
                rot[i] = Q[i] < n && jacobi_angle(A[P[i] * n + P[i]], A[Q[i] * n + Q[i]], A[P[i] * n + Q[i]], C[i], S[i]);
            }
            for (int i = 0; i < ne / 2; i++) {
                if (!rot[i]) continue;
                const int p = P[i], q = Q[i]; const float c = C[i], sn = S[i];
                for (int k = 0; k < n; k++) {      // columns p, q of A and of V
                    const float akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                    const float vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - sn * vkq;
                    V[k * n + q] = sn * vkp + c * vkq;
                }
            }
            for (int i = 0; i < ne / 2; i++) {
                if (!rot[i]) continue;
                const int p = P[i], q = Q[i]; const float c = C[i], sn = S[i];
                for (int k = 0; k < n; k++) {      // rows p, q
                    const float apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                }
            }
        }
    }
    // ascending by rank (ties keep index order), columns follow; A's storage receives the sorted vectors
    float d[16];
    for (int i = 0; i < n; i++) d[i] = A[i * n + i];
    for (int i = 0; i < n; i++) {
        int rank = 0;
        for (int j = 0; j < n; j++) if (d[j] < d[i] || (d[j] == d[i] && j < i)) rank++;
        w[rank] = d[i];
        for (int k = 0; k < n; k++) A[k * n + rank] = V[k * n + i];
    }
    for (int j = 0; j < n; j++) {
        float lead = 0.0f;
        for (int k = 0; k < n; k++) if (A[k * n + j] != 0.0f) { lead = A[k * n + j]; break; }
        for (int k = 0; k < n; k++) V[k * n + j] = (lead < 0.0f) ? -A[k * n + j] : A[k * n + j];
    }
}

// per eigen-direction part of ComputeGaussian (src/h2mc.cpp:88-118): `vtg` = (eigenvector . gradient)
LMC_HD void h2mc_eigen_scale(const Options &opt, float invSigmaSq, float w, float vtg, float &eigenBuff, float &offsetBuff, float &post) {
    float e = (dm_abs(w) > 1e-10f) ? 1.0f / dm_abs(w) : 0.0f;
    const float ob = e * vtg;
    float s2 = 1.0f, o = 0.0f;
    if (dm_abs(w) > 1e-10f) {
        o = ob;
        if (w > 0.0f) { s2 = opt.h2mcPosScale; o *= opt.h2mcPosOffset; }
        else { s2 = opt.h2mcNegScale; o *= opt.h2mcNegOffset; }
    } else {
        s2 = opt.h2mcL * opt.h2mcL;
        o = 0.5f * ob * opt.h2mcL * opt.h2mcL;
    }
    e *= s2;
    e = (e > 1e-10f) ? 1.0f / e : 0.0f;
    eigenBuff = e; offsetBuff = o; post = e + invSigmaSq;
}

// ComputeGaussian(h2mcParam, sc, vGrad, vHess, gaussian)  (src/h2mc.cpp:70-142 and :3-68).
// Diagonal outputs go to `g`, dense factors to covL / invCov; returns 1 when the result is dense.
template <int DIM>
LMC_HD_NOINLINE int h2mc_compute_gaussian(const Options &opt, float sc, int dim, const float *grad, float *hess,
                                          Gaussian<DIM> &g, float *covL, float *invCov) {
    const float sigma = opt.perturbStdDev;
    const float invSigmaSq = 1.0f / (sigma * sigma);
    g.dim = dim;
    float fp[16];                // Frobenius norm: lane l of 16 sums the entries l, l + 16, ..., then the tree
    for (int l = 0; l < 16; l++) { fp[l] = 0.0f; for (int i = l; i < dim * dim; i += 16) fp[l] += hess[i] * hess[i]; }
    const float frob = dm_sqrt(tree16(fp));
    if (sc <= 1e-15f || frob < 0.5f / (sigma * sigma)) {
        for (int i = 0; i < dim; i++) { g.mean[i] = 0.0f; g.covL_d[i] = sigma; g.invCov_d[i] = invSigmaSq; }
        g.logDet = 0.0f;
        for (int i = 0; i < dim; i++) g.logDet += dm_log(invSigmaSq);
        return 0;
    }
    float V[LMC_H2MC_DIM * LMC_H2MC_DIM], w[LMC_H2MC_DIM], eigenBuff[LMC_H2MC_DIM], offsetBuff[LMC_H2MC_DIM], post[LMC_H2MC_DIM];
    // SelfAdjointEigenSolver reads the lower triangle of the column-major map H(r, c) = vHess[c * dim + r]
    for (int r = 0; r < dim; r++) for (int c = 0; c < r; c++) hess[c * dim + r] = hess[r * dim + c];
    jacobi_eigen(dim, hess, V, w);
    for (int i = 0; i < dim; i++) {
        float vtg = 0.0f;
        for (int k = 0; k < dim; k++) vtg += V[k * dim + i] * grad[k];
        h2mc_eigen_scale(opt, invSigmaSq, w[i], vtg, eigenBuff[i], offsetBuff[i], post[i]);
    }
    for (int r = 0; r < dim; r++) {
        for (int c = 0; c < dim; c++) {
            float acc = 0.0f;
            for (int k = 0; k < dim; k++) acc += (V[r * dim + k] * post[k]) * V[c * dim + k];
            invCov[r * dim + c] = acc;
        }
        float m = 0.0f;
        for (int k = 0; k < dim; k++) m += V[r * dim + k] * ((eigenBuff[k] / post[k]) * offsetBuff[k]);
        g.mean[r] = m;
        for (int k = 0; k < dim; k++) covL[r * dim + k] = V[r * dim + k] * dm_sqrt(1.0f / post[k]);
        g.covL_d[r] = 0.0f; g.invCov_d[r] = 0.0f;
    }
    // sic: the template parameter `dim` is -1 for D > 12, so the reference's loop adds nothing
    g.logDet = 0.0f;
    if (dim <= 12) for (int i = 0; i < dim; i++) g.logDet += dm_log(post[i]);
    return 1;
}

// GaussianLogPdf / GenerateSample, dense branch (src/gaussian.cpp:29-31,49-51)
template <int DIM>
LMC_HD float gaussian_log_pdf_dense(const float *offset, float sign, const Gaussian<DIM> &g, const float *invCov) {
    float logPdf = (float)g.dim * (-0.9189385332046727f);
    logPdf += 0.5f * g.logDet;
    float q = 0.0f;
    for (int i = 0; i < g.dim; i++) {
        float r = 0.0f;
        for (int j = 0; j < g.dim; j++) r += invCov[i * g.dim + j] * (sign * offset[j] - g.mean[j]);
        q += (sign * offset[i] - g.mean[i]) * r;
    }
    logPdf -= 0.5f * q;
    return logPdf;
}
template <int DIM>
LMC_HD void generate_sample_dense(const Gaussian<DIM> &g, const float *covL, float *x, Rng &rng) {
    float z[DIM];
    NormalDist nd = normal_make(0.0f, 1.0f);
    for (int i = 0; i < g.dim; i++) z[i] = normal_draw(nd, rng);
    for (int i = 0; i < g.dim; i++) {
        float r = 0.0f;
        for (int j = 0; j < g.dim; j++) r += covL[i * g.dim + j] * z[j];
        x[i] = r + g.mean[i];
    }
}

// 0: no derivative function for this (c, l) -> IsotropicGaussian(sigma); 1: function exists but
// ssScore <= 1e-15 (zero gradient / Hessian); 2: evaluate gradient + Hessian (src/mutation_h2mc.h:62-90)
template <int MAXD>
LMC_HD int h2mc_grad_mode(const Scene &sc, const MarkovState<MAXD> &st) {
    const bool haveFunc = (st.sp.camDepth + st.sp.lightDepth - 1) <= sc.opt.maxDervDepth && grad_supported(sc, st.path) &&
                          path_dimension(st.path) <= LMC_H2MC_DIM;
    if (!haveFunc) return 0;
    return (st.sp.ssScore > 1e-15f) ? 2 : 1;
}

// initGaussian(state) of H2MCSmallStep::Mutate
template <int MAXD>
LMC_HD void h2mc_init_gaussian(const Scene &sc, MarkovState<MAXD> &st, int slot, int mode, const float *grad, H2mcSide *side,
                               bool built = false) {
    const int dim = path_dimension(st.path);
    if (built) {
        // the device's cooperative kernel (k_h2mc_gaussian) has already turned this state's gradient + Hessian into
        // its Gaussian -- same statements as h2mc_compute_gaussian, 16 lanes per matrix
    } else if (mode == 0) {
        isotropic_gaussian(dim, sc.opt.perturbStdDev, st.gaussian);
        side->dense[slot] = 0;
    } else {
        float g0[LMC_H2MC_DIM];
        for (int i = 0; i < dim; i++) g0[i] = (mode == 2) ? grad[i] : 0.0f;
        if (mode != 2) for (int i = 0; i < dim * dim; i++) side->hess[i] = 0.0f;
        side->dense[slot] = h2mc_compute_gaussian(sc.opt, st.sp.ssScore, dim, g0, side->hess, st.gaussian, side->covL[slot], side->invCov[slot]);
    }
    st.gaussianInitialized = 1;
}

template <int MAXD>
struct StepScratch {
    int kind;            // StepKind of the iteration in flight
    int needCurGrad;     // gradient of the CURRENT state wanted before the proposal (MALA, lazily)
    int needPropGrad;    // gradient of the PROPOSAL wanted before the acceptance test (MALA)
                         // (both: 2 = evaluated AND the H2MC Gaussian built from it, device only: k_h2mc_gaussian)
    int hasContrib;      // the proposal produced a contribution
    float a;             // acceptance probability (final after phase_finish)
    float offset[Limits<MAXD>::DIM];
    float grad[Limits<MAXD>::DIM];
};

// How MALASmallStep obtains the Gaussian of a state (src/mutation_mala.h:94-163 with the global
// cache never ready): 0 = IsotropicGaussian(malaStdDev); 1 = ComputeGaussian with a zero gradient
// (ssScore <= 1e-10, the reference skips dervFunc); 2 = ComputeGaussian with the evaluated gradient.
// With the global cache (opt.cacheEnabled) a fourth way: 3 = the dimension's cache is ready, no gradient is evaluated
// any more, the moments come from a cache query / are reused (src/mutation_mala.h:131-161).
LMC_HD bool cache_ready(const Scene &sc, int dim) {
    if (!sc.opt.cacheEnabled) return false;
    const int s = cache_slot(dim);
    return s >= 0 && sc.gc.ready[s] != 0;
}
template <int MAXD>
LMC_HD int mala_grad_mode(const Scene &sc, const MarkovState<MAXD> &st) {
    const int dim = path_dimension(st.path);
    const bool haveFunc = (st.sp.camDepth + st.sp.lightDepth - 1) <= sc.opt.maxDervDepth && grad_supported(sc, st.path);
    const bool inRange = dim >= sc.opt.pssMinLength && dim <= sc.opt.pssMaxLength;
    const bool ready = inRange && cache_ready(sc, dim);
    if (inRange && !ready && haveFunc) return (st.sp.ssScore > 1e-10f) ? 2 : 1;
    return ready ? 3 : 0;
}

// global_cache_t::query (src/global_cache.h:96-124): the entries within radius^2 = D * PSS_QUERY_DIST^2 of pss -- at
// most LMC_CACHE_KNN of them: the reference's patched nanoflann stops its KD-tree walk after 5 in-radius points
// (nanoflann.hpp:256-262), i.e. which 5 it keeps depends on the tree; here they are the first 5 in INSERTION order --
// averaged with weights 1 / (dist^2 + 1e-6), dist being the SQUARED distance nanoflann reports.
// the weighted average over the (at most LMC_CACHE_KNN) neighbours, in ascending insertion order
LMC_HD void cache_average(const float *base, int dim, int found, const int *idx, const float *dist, float *v1, float *v2) {
    const int stride = 3 * dim;
    double sum_w = 0.0;
    for (int i = 0; i < dim; i++) { v1[i] = 0.0f; v2[i] = 0.0f; }
    for (int k = 0; k < found; k++) {
        const float *q = base + (size_t)idx[k] * stride;
        const float w = 1.0f / (dist[k] * dist[k] + 1e-6f);
        for (int i = 0; i < dim; i++) { v1[i] += q[dim + i] * w; v2[i] += q[2 * dim + i] * w; }
        sum_w += (double)w;
    }
    for (int i = 0; i < dim; i++) { v1[i] = (float)((double)v1[i] / sum_w); v2[i] = (float)((double)v2[i] / sum_w); }
}
LMC_HD bool cache_reuse(int dim, int queried, const float *pss, const float *last_pss) {      // src/mutation_mala.h:222-232
    if (!queried) return false;
    float dist_sqr = 0.0f;
    for (int i = 0; i < dim; i++) { const float diff = pss[i] - last_pss[i]; dist_sqr += diff * diff; }
    return dist_sqr < (float)dim * (LMC_CACHE_REUSE_DIST * LMC_CACHE_REUSE_DIST);
}
LMC_HD bool cache_query(const Scene &sc, int dim, const float *pss, float *v1, float *v2) {
    const int s = cache_slot(dim);
    if (s < 0 || !sc.gc.ready[s]) return false;
    const float *base = sc.gc.data + cache_slot_offset(s);
    const int stride = 3 * dim;
    const float radius = (float)dim * (LMC_CACHE_QUERY_DIST * LMC_CACHE_QUERY_DIST);
    int idx[LMC_CACHE_KNN]; float dist[LMC_CACHE_KNN]; int found = 0;
    if (sc.gc.grid && sc.gc.gridReady[s]) {
        // the 27 cells around the query; keep the LMC_CACHE_KNN in-radius entries with the smallest insertion index,
        // in ascending order -- exactly what the linear scan below finds
        const int *cellStart = sc.gc.grid + (size_t)s * LMC_CACHE_GRID_INTS;
        const int *entry = cellStart + 2 * LMC_CACHE_CELLS + 1;
        const int c0 = cache_cell_coord(pss[0]), c1 = cache_cell_coord(pss[1]), c2 = cache_cell_coord(pss[2]);
        const int G = LMC_CACHE_GRID;
        for (int a = (c0 > 0 ? c0 - 1 : 0); a <= (c0 < G - 1 ? c0 + 1 : G - 1); a++)
            for (int b = (c1 > 0 ? c1 - 1 : 0); b <= (c1 < G - 1 ? c1 + 1 : G - 1); b++) {
                const int row = (a * G + b) * G;
                const int p0 = cellStart[row + (c2 > 0 ? c2 - 1 : 0)], p1 = cellStart[row + (c2 < G - 1 ? c2 + 1 : G - 1) + 1];
                for (int p = p0; p < p1; p++) {         // the three cells along the last axis are contiguous
                    const int e = entry[p];
                    const float *q = base + (size_t)e * stride;
                    float d = 0.0f;
                    for (int j = 0; j < dim; j++) { const float t = pss[j] - q[j]; d += t * t; }
                    if (!(d < radius)) continue;
                    int k = found < LMC_CACHE_KNN ? found : LMC_CACHE_KNN - 1;
                    if (found == LMC_CACHE_KNN && e > idx[k]) continue;
                    while (k > 0 && idx[k - 1] > e) { idx[k] = idx[k - 1]; dist[k] = dist[k - 1]; k--; }
                    idx[k] = e; dist[k] = d;
                    if (found < LMC_CACHE_KNN) found++;
                }
            }
    } else {
        for (int e = 0; e < LMC_CACHE_MAX_SIZE && found < LMC_CACHE_KNN; e++) {
            const float *q = base + (size_t)e * stride;
            float d = 0.0f;
            for (int j = 0; j < dim; j++) { const float t = pss[j] - q[j]; d += t * t; }
            if (d < radius) { idx[found] = e; dist[found] = d; found++; }
        }
    }
    if (!found) return false;
    cache_average(base, dim, found, idx, dist, v1, v2);
    return true;
}

// dervFunc + IsFinite guard (src/mutation_mala.h:100-110)
template <int MAXD>
LMC_HD void mala_eval_gradient(const Scene &sc, const MarkovState<MAXD> &st, float *grad, unsigned int *gradStats) {
    const int dim = path_dimension(st.path);
    path_gradient(sc, st.path, grad);
    bool finite = true;
    for (int i = 0; i < dim; i++) if (!dm_isfinite(grad[i])) finite = false;
    if (!finite) {
        for (int i = 0; i < dim; i++) grad[i] = 0.0f;
        if (gradStats) gradStats[1]++;
    }
    if (gradStats) gradStats[0]++;
}

// drift clamp, Adam-style moments, diagonal Gaussian (src/mutation_mala.h:111-129, src/mala.cpp:7-51).
// `grad` is read only when mode == 2.
template <int MAXD>
LMC_HD_NOINLINE void mala_finish_gaussian(const Scene &sc, const MarkovState<MAXD> &st, ChainVars<MAXD> &ch, int mode,
                                          const float *grad, float *new_v1, float *new_v2, Gaussian<Limits<MAXD>::DIM> &out) {
    const int dim = path_dimension(st.path);
    if (sc.opt.cacheEnabled) {      // GetPathPss(path, chain->pss); chain->pathWeight = lsScore (src/mutation_mala.h:89-92,184-187)
        get_path_pss(st.path, ch.pss);
        ch.pathWeight = st.sp.lsScore;
    }
    if (mode == 0) { isotropic_gaussian(dim, sc.opt.malaStdDev, out); return; }
    if (mode == 3) {                // src/mutation_mala.h:131-161 / 221-252
        const bool reuse = cache_reuse(dim, ch.queried, ch.pss, ch.last_pss);
        if (!reuse) {
            ch.cacheQueries += 1;
            bool hit;
            if (ch.preq) { hit = ch.preq == 2; ch.preq = 0; }       // answered ahead by k_cache_prequery (same statements, 32 lanes per query)
            else hit = cache_query(sc, dim, ch.pss, ch.v1, ch.v2);
            if (!hit) { isotropic_gaussian(dim, sc.opt.malaStdDev, out); return; }
            ch.cacheHits += 1;
            ch.queried = 1;
            for (int i = 0; i < dim; i++) ch.last_pss[i] = ch.pss[i];
        }
        float Mq[Limits<MAXD>::DIM];
        for (int i = 0; i < dim; i++) Mq[i] = dm_clamp(1.0f / (1e-3f + dm_sqrt(ch.v2[i])), LMC_PCD_MIN, LMC_PCD_MAX);
        compute_gaussian_lmc(dim, ch.v1, Mq, sc.opt.malaStepsize, sc.opt.malaStdDev, st.sp.ssScore, out);
        return;
    }
    float vGrad[Limits<MAXD>::DIM];
    for (int i = 0; i < dim; i++) vGrad[i] = (mode == 2) ? grad[i] : 0.0f;
    float norm = 0.0f;
    const float drift = sc.opt.malaGN;
    for (int i = 0; i < dim; i++) norm += vGrad[i] * vGrad[i];
    norm = dm_sqrt(norm);
    for (int i = 0; i < dim; i++) vGrad[i] *= drift / dm_max(drift, norm);
    bool first = true;                  // (no early exit: the loads of the record's vector stay independent)
    for (int i = 0; i < dim; i++) first &= !(new_v2[i] > 1e-10f);
    float M[Limits<MAXD>::DIM];
    for (int i = 0; i < dim; i++) {
        const float g = vGrad[i];
        new_v1[i] = first ? g : 0.9f * ch.v1[i] + 0.1f * g;
        new_v2[i] = first ? g * g : 0.999f * ch.v2[i] + 0.001f * g * g;
        M[i] = dm_clamp(1.0f / (1e-3f + dm_sqrt(new_v2[i])), LMC_PCD_MIN, LMC_PCD_MAX);
    }
    compute_gaussian_lmc(dim, new_v1, M, sc.opt.malaStepsize, sc.opt.malaStdDev, st.sp.ssScore, out);
}

struct RunParams {
    float normalization;
    int numChains;
    long long numSamplesThisChain;   // for the LS_RATIO large-step schedule (src/mlt.cpp:96)
    const float *initLsScore;        // initStates[i].spContrib.lsScore (outlier reset, src/mlt.cpp:152-158)
};

// Phase 0: choose the mutation (src/mlt.cpp:96-101, src/mutation_mala.h:47-81).
template <int MAXD>
LMC_HD void phase_begin(const Scene &sc, const RunParams &rp, long long sampleIdx, MarkovState<MAXD> &cur,
                        ChainVars<MAXD> &ch, Rng &rng, StepScratch<MAXD> &ss) {
    ss.needCurGrad = 0; ss.needPropGrad = 0; ss.hasContrib = 0; ss.a = 1.0f;
    if (sc.opt.cacheEnabled) ch.preq = 0;       // (cache runs only: the flag sits in a sector of the record nothing else touches here)
    const float lsScale = ((float)sampleIdx > (float)rp.numSamplesThisChain * sc.opt.lsRatio) ? sc.opt.largeStepProbScale : 1.0f;
    if (!cur.valid || rng_uniform(rng) < sc.opt.largeStepProbability * lsScale) { ss.kind = STEP_LARGE; return; }
    if (!sc.opt.mala && !sc.opt.h2mc) { ss.kind = STEP_ISO; return; }
    if (rng_uniform(rng) < sc.opt.uniformMixingProbability) { ss.kind = STEP_ISO; return; }
    if (sc.opt.h2mc) {          // src/mlt.cpp:71-74: h2mc takes precedence over mala
        ss.kind = STEP_H2MC;
        if (!cur.gaussianInitialized && h2mc_grad_mode(sc, cur) == 2) ss.needCurGrad = 1;
        return;
    }
    ss.kind = STEP_MALA;
    if (!ch.buffered) {
        for (int i = 0; i < Limits<MAXD>::DIM; i++) {
            ch.v1[i] = 0; ch.v2[i] = 0; ch.curr_new_v1[i] = 0; ch.curr_new_v2[i] = 0; ch.prop_new_v1[i] = 0; ch.prop_new_v2[i] = 0;
        }
        if (sc.opt.cacheEnabled) {
            for (int i = 0; i < Limits<MAXD>::DIM; i++) { ch.pss[i] = 0; ch.last_pss[i] = 0; }
            ch.queried = 0;
        }
        ch.buffered = 1;
    }
    if (!cur.gaussianInitialized && mala_grad_mode(sc, cur) == 2) ss.needCurGrad = 1;
}

// Phase 1 / 3: PSS gradient of the current state / of the proposal (no RNG).
// ORDER: -1 = decided at run time (host twin); 1 = gradient only (LMC); 2 = gradient + Hessian (H2MC)
template <int MAXD, int ORDER = -1>
LMC_HD void phase_gradient(const Scene &sc, const MarkovState<MAXD> &st, StepScratch<MAXD> &ss, unsigned int *gradStats,
                           H2mcSide *side) {
    if (ORDER == 1 && ss.kind == STEP_H2MC) return;
    if (ORDER == 2 && ss.kind != STEP_H2MC) return;
    if (ORDER != 1 && ss.kind == STEP_H2MC) {
        // gradient + Hessian, IsFinite guard on both (src/mutation_h2mc.h:74-86)
        const int dim = path_dimension(st.path);
        path_hessian(sc, st.path, ss.grad, side->hess);
        bool finite = true;
        for (int i = 0; i < dim; i++) if (!dm_isfinite(ss.grad[i])) finite = false;
        for (int i = 0; i < dim * dim; i++) if (!dm_isfinite(side->hess[i])) finite = false;
        if (!finite) {
            for (int i = 0; i < dim; i++) ss.grad[i] = 0.0f;
            for (int i = 0; i < dim * dim; i++) side->hess[i] = 0.0f;
            if (gradStats) gradStats[1]++;
        }
        if (gradStats) gradStats[0]++;
        return;
    }
    if (ORDER == 2) return;
    mala_eval_gradient(sc, st, ss.grad, gradStats);
}

// Phase 2: draw the proposal and trace it.
// SmallStep::Mutate (src/mutation_small.h:16-55), MALASmallStep::Mutate up to the proposal's
// gradient (src/mutation_mala.h:83-176), LargeStep::Mutate (src/mutation_large.h:31-127).
// Each mutation is cut around its path-tracing call into a PRE part (everything before
// GeneratePathBidir / PerturbPathBidir) and a POST part (everything after), so the wavefront can
// run the tracing in between as per-vertex stages (stages.h) while the host twin calls the
// monolithic path functions.

// LargeStep::Mutate before GeneratePathBidir
template <int MAXD>
LMC_HD void propose_pre_large(MarkovState<MAXD> &prop, ChainVars<MAXD> &ch) {
    ch.lastMutationType = MUT_LARGE;
    path_clear(prop.path);
}
// LargeStep::Mutate after GeneratePathBidir (src/mutation_large.h:60-127): contribution choice, acceptance, splats
template <int MAXD, class CL>
LMC_HD void propose_post_large(const RunParams &rp, const MarkovState<MAXD> &cur, MarkovState<MAXD> &prop,
                               const ChainVars<MAXD> &ch, Rng &rng, StepScratch<MAXD> &ss, const CL &contribs) {
    float a = 1.0f;
    prop.gaussianInitialized = 0;
    if (contribs.n > 0) {
        float cdf[Limits<MAXD>::MAXC + 1];
        cdf[0] = 0.0f;
        for (int i = 0; i < contribs.n; i++) cdf[i + 1] = cdf[i] + contribs.c[i].lsScore;
        const float scoreSum = cdf[contribs.n];
        const float invSc = inverse(scoreSum);
        for (int i = 0; i <= contribs.n; i++) cdf[i] *= invSc;
        const int it = upper_bound_f(cdf, contribs.n + 1, rng_uniform(rng));
        const int contribId = dm_clampi(it - 1, 0, contribs.n - 1);
        prop.sp = contribs.c[contribId];
        prop.scoreSum = scoreSum;
        if (cur.valid) {
            const float probProposal = prop.sp.lsScore / prop.scoreSum;
            const float probLast = ch.lastScore / ch.lastScoreSum;
            a = dm_clamp((prop.sp.lsScore * probLast) / (cur.sp.lsScore * probProposal), 0.0f, 1.0f);
        }
        prop.nSplat = contribs.n;
        for (int i = 0; i < contribs.n; i++) {
            prop.splat[i].screenPos = contribs.c[i].screenPos;
            prop.splat[i].contrib = contribs.c[i].contrib * (rp.normalization / scoreSum);
        }
    } else {
        a = 0.0f;
    }
    ss.a = a;
}

// small steps before PerturbPathBidir: the proposal offset in ss.offset, prop.path = cur.path.
// COPY_PATH = false leaves the path alone: the device wavefront carries the PathHead in its payload and
// reads each vertex from cur.path / writes it to prop.path as PerturbPathBidir reaches it.
template <int MAXD, bool COPY_PATH = true>
LMC_HD void propose_pre_small(const Scene &sc, MarkovState<MAXD> &cur, MarkovState<MAXD> &prop, ChainVars<MAXD> &ch,
                              Rng &rng, StepScratch<MAXD> &ss, H2mcSide *side, int curSlot) {
    const int dim = path_dimension(cur.path);
    if (ss.kind == STEP_ISO) {
        if (COPY_PATH) path_copy(prop.path, cur.path);
        NormalDist nd = normal_make(0.0f, sc.opt.perturbStdDev);
        ch.lastMutationType = MUT_SMALL;
        for (int i = 0; i < dim; i++) ss.offset[i] = normal_draw(nd, rng);
        return;
    }
    if (ss.kind == STEP_H2MC) {
        ch.lastMutationType = MUT_H2MC_SMALL;
        if (!cur.gaussianInitialized) h2mc_init_gaussian(sc, cur, curSlot, h2mc_grad_mode(sc, cur), ss.grad, side, ss.needCurGrad == 2);
        if (side->dense[curSlot]) generate_sample_dense(cur.gaussian, side->covL[curSlot], ss.offset, rng);
        else generate_sample(cur.gaussian, ss.offset, rng);
        if (COPY_PATH) path_copy(prop.path, cur.path);
        return;
    }
    // STEP_MALA
    ch.lastMutationType = MUT_MALA_SMALL;
    if (!cur.gaussianInitialized) {
        mala_finish_gaussian(sc, cur, ch, mala_grad_mode(sc, cur), ss.grad, ch.curr_new_v1, ch.curr_new_v2, cur.gaussian);
        cur.gaussianInitialized = 1;
    }
    generate_sample(cur.gaussian, ss.offset, rng);
    if (COPY_PATH) path_copy(prop.path, cur.path);
}
// small steps after PerturbPathBidir (contribs holds 0 or 1 entries)
template <int MAXD, class CL>
LMC_HD void propose_post_small(const Scene &sc, const RunParams &rp, const MarkovState<MAXD> &cur, MarkovState<MAXD> &prop,
                               StepScratch<MAXD> &ss, const CL &contribs) {
    if (ss.kind == STEP_ISO) {
        prop.gaussianInitialized = 0;
        if (contribs.n > 0) {
            prop.sp = contribs.c[0];
            ss.a = dm_clamp(prop.sp.ssScore / cur.sp.ssScore, 0.0f, 1.0f);
            prop.nSplat = 1;
            prop.splat[0].screenPos = prop.sp.screenPos;
            prop.splat[0].contrib = prop.sp.contrib * (rp.normalization / prop.sp.lsScore);
        } else {
            ss.a = 0.0f;
        }
        return;
    }
    if (contribs.n > 0) {
        prop.sp = contribs.c[0];
        ss.hasContrib = 1;
        if (ss.kind == STEP_H2MC) { if (h2mc_grad_mode(sc, prop) == 2) ss.needPropGrad = 1; }
        else if (mala_grad_mode(sc, prop) == 2) ss.needPropGrad = 1;
    } else {
        ss.a = 0.0f;
    }
}

// ONLY: -1 = any kind (host twin); 0 = large steps only; 1 = small steps only
template <int MAXD, int ONLY = -1>
LMC_HD void phase_propose(const Scene &sc, const RunParams &rp, MarkovState<MAXD> &cur, MarkovState<MAXD> &prop,
                          ChainVars<MAXD> &ch, Rng &rng, StepScratch<MAXD> &ss, H2mcSide *side, int curSlot) {
    if (ONLY == 1 && ss.kind == STEP_LARGE) return;
    if (ONLY == 0 && ss.kind != STEP_LARGE) return;
    if (ONLY != 1 && ss.kind == STEP_LARGE) {
        propose_pre_large(prop, ch);
        ContribList<Limits<MAXD>::MAXC> contribs; contribs.clear();
        const int minDepth = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
        generate_path_bidir(sc, minDepth, sc.opt.maxDepth, prop.path, contribs, rng);
        propose_post_large(rp, cur, prop, ch, rng, ss, contribs);
        return;
    }
    if (ONLY == 0) return;
    ContribList<2> contribs; contribs.clear();
    propose_pre_small(sc, cur, prop, ch, rng, ss, side, curSlot);
    perturb_path_bidir(sc, ss.offset, prop.path, contribs, rng);
    propose_post_small(sc, rp, cur, prop, ss, contribs);
}

// The same phase through the staged path functions (stages.h), ray queries answered on the spot:
// the host-side model of the device wavefront (tests/test_staged.py compares it with phase_propose).
struct ImmediateShadowSink {
    static const bool kDeferConnections = false;
    const Scene *sc;
    template <class LS>
    LMC_HD void emit_connection(const Scene &, int, int, int, const LS *, const SurfaceVertex *, const LS &, const SurfaceVertex &, V2,
                                SubpathContrib *, int *) {}
    LMC_HD void emit(const Ray &ray, float dist, int, int *flag) { *flag = cand_resolve(*flag, scene_occluded(*sc, ray, dist)); }
};
template <int MAXD>
struct StagedWork { TraceState ts; PropCand pc; GenWork<MAXD, Limits<MAXD>::MAXC> gw; };

template <int MAXD>
LMC_HD void phase_propose_staged(const Scene &sc, const RunParams &rp, MarkovState<MAXD> &cur, MarkovState<MAXD> &prop,
                                 ChainVars<MAXD> &ch, Rng &rng, StepScratch<MAXD> &ss, H2mcSide *side, int curSlot,
                                 StagedWork<MAXD> &w) {
    ImmediateShadowSink sink; sink.sc = &sc;
    DeferredList<ImmediateShadowSink> dl;
    TraceState &ts = w.ts;
    Path<MAXD> &path = prop.path;
    if (ss.kind == STEP_LARGE) {
        propose_pre_large(prop, ch);
        w.gw.n = 0;
        dl.bind(w.gw.c, w.gw.flag, &w.gw.n, Limits<MAXD>::MAXC, &sink);
        const int minDepth = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
        bool more = gen_stage_begin(sc, path, ts, w.gw.ls, rng);
        while (more) {
            const Hit h = bvh_traverse<false>(sc, ts.ray, ts.minT, ts.maxT);
            if (ts.stage == TS_G_LGT) more = gen_stage_light(sc, minDepth, sc.opt.maxDepth, path, path.lgt[path.nLgt], ts, w.gw.ls, dl, rng, h);
            else more = gen_stage_camera(sc, minDepth, sc.opt.maxDepth, path, path.cam[path.nCam], path.lgt, ts, w.gw.ls, dl, rng, h);
        }
        w.gw.n = deferred_compact(w.gw.c, w.gw.flag, w.gw.n);
        propose_post_large(rp, cur, prop, ch, rng, ss, w.gw);
        return;
    }
    propose_pre_small(sc, cur, prop, ch, rng, ss, side, curSlot);
    w.pc.n = 0;
    dl.bind(w.pc.c, w.pc.flag, &w.pc.n, 2, &sink);
    const float *offset = ss.offset;
    bool more = perturb_stage_begin(sc, offset, path, ts, rng);
    while (more) {
        const Hit h = bvh_traverse<false>(sc, ts.ray, ts.minT, ts.maxT);
        if (ts.stage == TS_P_LGT) more = perturb_stage_light(sc, offset, path, path.lgt[ts.depth], ts, dl, rng, h);
        else more = perturb_stage_camera(sc, offset, path, path.cam[ts.depth], path.lgt, ts, dl, rng, h);
    }
    w.pc.n = deferred_compact(w.pc.c, w.pc.flag, w.pc.n);
    propose_post_small(sc, rp, cur, prop, ss, w.pc);
}

struct StepInfo {          // what one iteration did (parity traces / stats)
    int mutationType;
    int accepted;
    float a;
};

// Phase 4: proposal Gaussian + MH ratio (src/mutation_mala.h:178-272), splats, acceptance test,
// state swap, moment commit, outlier reset (src/mlt.cpp:103-170).
template <int MAXD, class FILM>
LMC_HD StepInfo phase_finish(const Scene &sc, const RunParams &rp, int chainId, long long sampleIdx,
                             MarkovState<MAXD> *states, int &curIdx, ChainVars<MAXD> &ch, Rng &rng, FILM &film,
                             StepScratch<MAXD> &ss, H2mcSide *side) {
    MarkovState<MAXD> &cur = states[curIdx];
    MarkovState<MAXD> &prop = states[curIdx ^ 1];
    if (ss.kind == STEP_H2MC && ss.hasContrib) {
        const int cs_ = curIdx, ps_ = curIdx ^ 1;
        h2mc_init_gaussian(sc, prop, ps_, h2mc_grad_mode(sc, prop), ss.grad, side, ss.needPropGrad == 2);
        const float py = side->dense[cs_] ? gaussian_log_pdf_dense(ss.offset, 1.0f, cur.gaussian, side->invCov[cs_])
                                          : gaussian_log_pdf(ss.offset, 1.0f, cur.gaussian);
        const float px = side->dense[ps_] ? gaussian_log_pdf_dense(ss.offset, -1.0f, prop.gaussian, side->invCov[ps_])
                                          : gaussian_log_pdf(ss.offset, -1.0f, prop.gaussian);
        ss.a = dm_clamp(dm_exp(px - py) * prop.sp.ssScore / cur.sp.ssScore, 0.0f, 1.0f);
        prop.nSplat = 1;
        prop.splat[0].screenPos = prop.sp.screenPos;
        prop.splat[0].contrib = prop.sp.contrib * (rp.normalization / prop.sp.lsScore);
    }
    if (ss.kind == STEP_MALA && ss.hasContrib) {
        mala_finish_gaussian(sc, prop, ch, mala_grad_mode(sc, prop), ss.grad, ch.prop_new_v1, ch.prop_new_v2, prop.gaussian);
        prop.gaussianInitialized = 1;
        const float py = gaussian_log_pdf(ss.offset, 1.0f, cur.gaussian);
        const float px = gaussian_log_pdf(ss.offset, -1.0f, prop.gaussian);
        ss.a = dm_clamp(dm_exp(px - py) * prop.sp.ssScore / cur.sp.ssScore, 0.0f, 1.0f);
        prop.nSplat = 1;
        prop.splat[0].screenPos = prop.sp.screenPos;
        prop.splat[0].contrib = prop.sp.contrib * rp.normalization / prop.sp.lsScore;
    }
    const float a = ss.a;
    const bool isLargeStep = ss.kind == STEP_LARGE;
    const int W = sc.cam.width, H = sc.cam.height;
    if (cur.valid && a < 1.0f) {
        for (int i = 0; i < cur.nSplat; i++) splat(film, W, H, cur.splat[i].screenPos, (1.0f - a) * cur.splat[i].contrib);
    }
    if (a > 0.0f) {
        for (int i = 0; i < prop.nSplat; i++) splat(film, W, H, prop.splat[i].screenPos, a * prop.splat[i].contrib);
    }
    StepInfo info; info.mutationType = ch.lastMutationType; info.accepted = 0; info.a = a;
    if (a > 0.0f && rng_uniform(rng) <= a) {
        info.accepted = 1;
        to_subpath(prop.sp.camDepth, prop.sp.lightDepth, prop.path);
        curIdx ^= 1;
        MarkovState<MAXD> &ncur = states[curIdx];
        ncur.valid = 1;
        ch.adjacentReject = 0;
        if (isLargeStep) {
            if (sc.opt.cacheEnabled && ch.buffered && ch.pathWeight > 1e-10f) {      // src/mlt.cpp:121-127
                const int pdim = path_dimension(states[curIdx ^ 1].path);             // the state the chain leaves
                if (pdim >= sc.opt.pssMinLength && pdim <= sc.opt.pssMaxLength && cache_slot(pdim) >= 0 && !cache_ready(sc, pdim))
                    ch.pushDim = pdim;        // committed in chain order after the iteration (cache_commit / k_cache_*)
            }
            ch.lastScoreSum = ncur.scoreSum;
            ch.lastScore = ncur.sp.lsScore;
            ncur.gaussianInitialized = 0;
            ch.buffered = 0;
        } else if (ch.lastMutationType == MUT_MALA_SMALL) {
            for (int i = 0; i < Limits<MAXD>::DIM; i++) { ch.v1[i] = ch.prop_new_v1[i]; ch.v2[i] = ch.prop_new_v2[i]; }
            ch.t += 1;
            ch.buffered = 1;
            ncur.gaussianInitialized = 1;
        }
    } else {
        ch.adjacentReject += 1;
        const bool strongReject = cur.sp.lsScore > sc.opt.outlierRatioThreshold * rp.normalization;
        if (ch.adjacentReject > sc.opt.outlierWeakRejectCnt ||
            (strongReject && ch.adjacentReject > sc.opt.outlierStrongRejectCnt)) {
            int cid = chainId, cnt = 0;
            for (;;) {
                cur.sp.lsScore = rp.initLsScore[cid];
                if (cur.sp.lsScore < sc.opt.outlierRatioThreshold * rp.normalization) break;
                cid = (int)(((long long)cid + sampleIdx + (long long)(cnt++)) % (long long)rp.numChains);
                if (cnt > rp.numChains) break;   // guard: the reference would spin forever
            }
            cur.valid = 0; cur.gaussianInitialized = 0; cur.nSplat = 0;
            prop.valid = 0; prop.gaussianInitialized = 0; prop.nSplat = 0;
            path_clear(prop.path);
            ch.buffered = 0;
            ch.outlierResets += 1;
        }
    }
    return info;
}

// One whole iteration of the loop at src/mlt.cpp:91-170 (host twin; the device runs the phases
// as separate kernels).  `cur` and `prop` are swapped by index.
template <int MAXD, class FILM>
LMC_HD StepInfo chain_step(const Scene &sc, const RunParams &rp, int chainId, long long sampleIdx,
                           MarkovState<MAXD> *states, int &curIdx, ChainVars<MAXD> &ch, Rng &rng, FILM &film,
                           unsigned int *gradStats, StepScratch<MAXD> &ss, H2mcSide *side, StagedWork<MAXD> *staged = nullptr) {
    phase_begin(sc, rp, sampleIdx, states[curIdx], ch, rng, ss);
    if (ss.needCurGrad) phase_gradient(sc, states[curIdx], ss, gradStats, side);
    if (staged) phase_propose_staged(sc, rp, states[curIdx], states[curIdx ^ 1], ch, rng, ss, side, curIdx, *staged);
    else phase_propose(sc, rp, states[curIdx], states[curIdx ^ 1], ch, rng, ss, side, curIdx);
    if (ss.needPropGrad) phase_gradient(sc, states[curIdx ^ 1], ss, gradStats, side);
    return phase_finish(sc, rp, chainId, sampleIdx, states, curIdx, ch, rng, film, ss, side);
}

}  // namespace lmc
