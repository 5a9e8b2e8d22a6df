// h2mc_kernels.cuh -- warp-cooperative ComputeGaussian of the H2MC mutation (src/h2mc.cpp:3-142).
//
// The reference turns a state's PSS gradient + Hessian into an anisotropic Gaussian per chain with Eigen's
// SelfAdjointEigenSolver on a <= 16 x 16 matrix, one chain per CPU thread.  One CUDA thread per chain is the wrong shape
// for that on a GPU: the matrices differ in size and in the number of sweeps they need, only every fifth chain needs a
// solve at all, and 3 x 256 floats of per-thread scratch live in local memory (measured: 1.6 of 32 lanes active,
// 47 ms of a 115 ms iteration at 2^20 chains).  Here a HALF-WARP owns one matrix:
//   * the chains that need a Gaussian are exactly the entries of the class-sorted gradient list the Hessian kernel
//     (k_wave_grad<2>) has just served, so this kernel walks the same list -- 16 lanes per entry;
//   * A and V sit in shared memory with row stride 17 (rows and columns are both conflict-free);
//   * the Jacobi rotations run in the parallel (round-robin) order of core/mutation.h: the n / 2 disjoint pairs of a
//     round get their angles from lanes 0 .. n/2-1, then lane l applies all of them to row l (column phase, A and V)
//     and to column l (row phase) -- no index arithmetic in the hot loops, no bank conflicts;
//   * norms, convergence tests and dot products are xor-shuffle trees over the 16 lanes (group_tree16 == tree16).
// Statement for statement this is h2mc_compute_gaussian + jacobi_eigen of core/mutation.h, so the result is bit-identical
// to the host twin (tests/test_h2mc.py); the chain kernels find the Gaussian built (StepScratch::need*Grad == 2).
#pragma once

namespace lmc_cuda {
using namespace lmc;

#define LMC_H2MC_BLOCK 128                 // 8 matrices per block
#define LMC_H2MC_LD 17

struct H2mcShared {
    float A[16 * LMC_H2MC_LD];
    float V[16 * LMC_H2MC_LD];
    float c[8], s[8];
    int p[8], q[8];                        // p < 0: this pair has nothing to rotate (A[p][q] == 0)
    float w[16], eb[16], ob[16], post[16], lg[16], grad[16];
};

__device__ __forceinline__ float group_tree16(float x, unsigned mask) {
    for (int off = 8; off > 0; off >>= 1) x += __shfl_xor_sync(mask, x, off);
    return x;
}

template <int MAXD>
__global__ void __launch_bounds__(LMC_H2MC_BLOCK) k_h2mc_gaussian(const __grid_constant__ Scene sc, ChainRec<MAXD> *states,
                                                                  const int *list, const int *count, int which, H2mcSide *sides) {
    __shared__ H2mcShared shm[LMC_H2MC_BLOCK / 16];
    const int grp = threadIdx.x >> 4, l = threadIdx.x & 15;
    const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
    H2mcShared &S = shm[grp];
    const int LD = LMC_H2MC_LD;
    const int total = *count;
    const float sigma = sc.opt.perturbStdDev;
    const float invSigmaSq = 1.0f / (sigma * sigma);
    for (int e = blockIdx.x * (LMC_H2MC_BLOCK / 16) + grp; e < total; e += gridDim.x * (LMC_H2MC_BLOCK / 16)) {
        const int ci = list[e];
        if (ci < 0) continue;                                   // alignment gap of the class-sorted list (group-uniform)
        ChainState<MAXD> &cs = states[ci].cs;
        if (cs.ss.kind != STEP_H2MC) continue;
        const int slot = cs.curIdx ^ which;
        MarkovState<MAXD> &st = cs.st[slot];
        H2mcSide *side = sides + ci;
        const int n = path_dimension(st.path);
        Gaussian<Limits<MAXD>::DIM> &g = st.gaussian;
        // ---- load; Frobenius norm of the Hessian as delivered
        S.grad[l] = (l < n) ? cs.ss.grad[l] : 0.0f;
        float fp = 0.0f;
        for (int idx = l; idx < n * n; idx += 16) {
            const float h = side->hess[idx];
            fp += h * h;
            S.A[(idx / n) * LD + (idx % n)] = h;
        }
        const float frob = dm_sqrt(group_tree16(fp, mask));
        __syncwarp(mask);
        if (st.sp.ssScore <= 1e-15f || frob < 0.5f / (sigma * sigma)) {
            if (l < n) { g.mean[l] = 0.0f; g.covL_d[l] = sigma; g.invCov_d[l] = invSigmaSq; }
            if (l == 0) {
                const float lg = dm_log(invSigmaSq);
                float logDet = 0.0f;
                for (int i = 0; i < n; i++) logDet += lg;
                g.logDet = logDet; g.dim = n;
                side->dense[slot] = 0;
            }
        } else {
            // the solver reads the lower triangle (column-major map of the reference, src/h2mc.cpp:77)
            for (int idx = l; idx < n * n; idx += 16) {
                const int r = idx / n, c = idx % n;
                if (c < r) S.A[c * LD + r] = S.A[r * LD + c];
                S.V[r * LD + c] = (r == c) ? 1.0f : 0.0f;
            }
            __syncwarp(mask);
            for (int sweep = 0; sweep < LMC_JACOBI_SWEEPS; sweep++) {
                float diagp = 0.0f, offp = 0.0f;
                if (l < n) {
                    const float d = S.A[l * LD + l];
                    diagp = d * d;
                    for (int q = l + 1; q < n; q++) { const float a = S.A[l * LD + q]; offp += a * a; }
                }
                const float off = group_tree16(offp, mask), diag = group_tree16(diagp, mask);
                if (jacobi_converged(off, diag)) break;
                for (int r = 0; r < n - 1; r++) {
                    if (l < n / 2) {
                        int p, q; float c, sn;
                        jacobi_pair(n, r, l, p, q);
                        const bool rot = jacobi_angle(S.A[p * LD + p], S.A[q * LD + q], S.A[p * LD + q], c, sn);
                        S.p[l] = rot ? p : -1; S.q[l] = q; S.c[l] = c; S.s[l] = sn;
                    }
                    __syncwarp(mask);
                    // lane l owns row l in the column phase and column l in the row phase (both conflict-free with the
                    // stride-17 rows) and walks the round's pairs: no index arithmetic in the hot loops
                    if (l < n) {
                        for (int i = 0; i < n / 2; i++) {                        // columns p, q of A and of V, row l
                            const int p = S.p[i];
                            if (p < 0) continue;
                            const int q = S.q[i]; const float c = S.c[i], sn = S.s[i];
                            const float akp = S.A[l * LD + p], akq = S.A[l * LD + q];
                            S.A[l * LD + p] = c * akp - sn * akq;
                            S.A[l * LD + q] = sn * akp + c * akq;
                            const float vkp = S.V[l * LD + p], vkq = S.V[l * LD + q];
                            S.V[l * LD + p] = c * vkp - sn * vkq;
                            S.V[l * LD + q] = sn * vkp + c * vkq;
                        }
                    }
                    __syncwarp(mask);
                    if (l < n) {
                        for (int i = 0; i < n / 2; i++) {                        // rows p, q, column l
                            const int p = S.p[i];
                            if (p < 0) continue;
                            const int q = S.q[i]; const float c = S.c[i], sn = S.s[i];
                            const float apk = S.A[p * LD + l], aqk = S.A[q * LD + l];
                            S.A[p * LD + l] = c * apk - sn * aqk;
                            S.A[q * LD + l] = sn * apk + c * aqk;
                        }
                    }
                    __syncwarp(mask);
                }
            }
            // ---- ascending by rank, columns follow (A's storage receives the sorted vectors), sign convention
            if (l < n) S.eb[l] = S.A[l * LD + l];
            __syncwarp(mask);
            if (l < n) {
                const float d = S.eb[l];
                int rank = 0;
                for (int j = 0; j < n; j++) { const float dj = S.eb[j]; if (dj < d || (dj == d && j < l)) rank++; }
                S.w[rank] = d;
                for (int k = 0; k < n; k++) S.A[k * LD + rank] = S.V[k * LD + l];
            }
            __syncwarp(mask);
            if (l < n) {
                float lead = 0.0f;
                for (int k = 0; k < n; k++) { const float v = S.A[k * LD + l]; if (v != 0.0f) { lead = v; break; } }
                for (int k = 0; k < n; k++) { const float v = S.A[k * LD + l]; S.V[k * LD + l] = (lead < 0.0f) ? -v : v; }
            }
            __syncwarp(mask);
            // ---- per eigen-direction scaling, then the dense factors
            if (l < n) {
                float vtg = 0.0f;
                for (int k = 0; k < n; k++) vtg += S.V[k * LD + l] * S.grad[k];
                float eb, ob, post;
                h2mc_eigen_scale(sc.opt, invSigmaSq, S.w[l], vtg, eb, ob, post);
                S.eb[l] = eb; S.ob[l] = ob; S.post[l] = post;
                S.lg[l] = (n <= 12) ? dm_log(post) : 0.0f;
            }
            __syncwarp(mask);
            float *invCov = side->invCov[slot], *covL = side->covL[slot];
            for (int idx = l; idx < n * n; idx += 16) {
                const int r = idx / n, c = idx % n;
                float acc = 0.0f;
                for (int k = 0; k < n; k++) acc += (S.V[r * LD + k] * S.post[k]) * S.V[c * LD + k];
                invCov[idx] = acc;
                covL[idx] = S.V[r * LD + c] * dm_sqrt(1.0f / S.post[c]);
            }
            if (l < n) {
                float m = 0.0f;
                for (int k = 0; k < n; k++) m += S.V[l * LD + k] * ((S.eb[k] / S.post[k]) * S.ob[k]);
                g.mean[l] = m; g.covL_d[l] = 0.0f; g.invCov_d[l] = 0.0f;
            }
            if (l == 0) {
                // sic: the reference's loop adds nothing for D > 12
                float logDet = 0.0f;
                if (n <= 12) for (int i = 0; i < n; i++) logDet += S.lg[i];
                g.logDet = logDet; g.dim = n;
                side->dense[slot] = 1;
            }
        }
        if (l == 0) { if (which) cs.ss.needPropGrad = 2; else cs.ss.needCurGrad = 2; }
        __syncwarp(mask);                                       // the group's shared slot is reused by its next entry
    }
}

}  // namespace lmc_cuda
