// mlt_init.h -- host restatement of MLTInit (reference: src/mlt.h:41-154).
//
// MLTInit stays on the host ("before it: DirectLighting, MLTInit (host)").
// What the chain loop actually consumes from it:
//   * normalization = avgScore = sum(lsScore) / numInitSamples            (src/mlt.h:150-153)
//   * initStates[i].spContrib.lsScore -- read only by the outlier reset   (src/mlt.cpp:152-158)
// The init states themselves start `valid = false` (MarkovState{false}, src/mlt.h:124), so every
// chain's first iteration is a large step and the seeded path is never used; the replay of
// src/mlt.h:124-146 therefore reduces to "take the lsScore of the selected init sample" (a
// deterministic replay reproduces the same contribution).
//
// Determinism: the reference seeds one RNG per hardware thread (RNG(threadId + seedOffset),
// NumSystemCores() threads, src/mlt.h:51-52,67), which makes the result machine dependent;
// here the number of LOGICAL threads is a parameter (default 32 = the reference machine) and
// logical thread t always generates the same samples wherever it runs.  Contributions are
// appended in (logical thread, sample, contribution) order; the reference's order depends on
// the mutex race.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <thread>
#include <vector>
#include "../core/mutation.h"

namespace lmc_host {

struct InitResult {
    float normalization;
    std::vector<float> initLsScore;     // per chain
    std::vector<float> lengthContrib;   // lsScore mass per path length (lengthDist input)
    long long numInitContribs;
};

// The sequential part of MLTInit: score sum, CDF and equal-spaced seeding (src/mlt.h:107-153) over the
// lsScores of all init contributions in (logical thread, sample, contribution) order.  Shared by the host
// path generator below and the device one (lmc_mlt_init_device: the kernel produces `scores`, the fp32
// running sums stay sequential so both give the same bits).
inline void mlt_init_finish(const std::vector<float> &scores, long long numInitSamples, int numChains, InitResult &res) {
    using namespace lmc;
    float totalScore = 0.0f;
    for (float s : scores) totalScore += s;
    res.numInitContribs = (long long)scores.size();
    if ((long long)scores.size() < (long long)numChains)
        throw std::runtime_error("MLT initialization failed, consider using a larger number of initial samples or smaller number of chains");

    // Equal-spaced seeding (src/mlt.h:107-148)
    std::vector<float> cdf(scores.size() + 1);
    cdf[0] = 0.0f;
    for (size_t i = 0; i < scores.size(); i++) cdf[i + 1] = cdf[i] + scores[i];
    const float interval = cdf.back() / (float)numChains;
    uint32_t tab[64];
    Rng rng; rng.tab = tab; rng.stride = 1;
    rng_seed(rng, (uint64_t)scores.size());
    float pos = rng_uniform_ab(rng, 0.0f, interval);
    int cdfPos = 0;
    res.initLsScore.resize(numChains);
    for (int i = 0; i < numChains; i++) {
        while (pos > cdf[cdfPos]) {
            if (cdfPos == (int)scores.size() - 1) break;   // guard: the reference would spin forever here
            cdfPos = std::min(cdfPos + 1, (int)scores.size() - 1);
        }
        // sic: mStates[cdfPos - 1]; cdfPos >= 1 whenever pos > 0
        res.initLsScore[i] = scores[cdfPos > 0 ? cdfPos - 1 : 0];
        pos += interval;
    }
    res.normalization = totalScore * (1.0f / (float)numInitSamples);
}

template <int MAXD>
inline void mlt_init(const lmc::Scene &sc, long long numInitSamples, int numChains, int logicalThreads,
                     InitResult &res) {
    using namespace lmc;
    if (logicalThreads < 1) logicalThreads = 1;
    struct LightState { float lsScore; };
    std::vector<std::vector<float>> perThread(logicalThreads);
    std::vector<std::vector<float>> perThreadLen(logicalThreads);
    const long long perT = numInitSamples / logicalThreads;
    const long long extra = numInitSamples % logicalThreads;
    const int minPathLength = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
    auto work = [&](int t) {
        uint32_t tab[64];
        Rng rng; rng.tab = tab; rng.stride = 1;
        rng_seed(rng, (uint64_t)(long long)(t + sc.opt.seedOffset));
        // sic: the reference tests `threadIndex` (a thread-local pool index) here; with one
        // logical thread per pool thread that is the same number.
        const long long n = perT + ((t < extra) ? 1 : 0);
        Path<MAXD> *path = new Path<MAXD>();
        ContribList<Limits<MAXD>::MAXC> contribs;
        std::vector<float> &lenC = perThreadLen[t];
        for (long long s = 0; s < n; s++) {
            contribs.clear();
            path_clear(*path);
            generate_path_bidir(sc, minPathLength, sc.opt.maxDepth, *path, contribs, rng);
            for (int i = 0; i < contribs.n; i++) {
                const SubpathContrib &c = contribs.c[i];
                const int len = c.camDepth + c.lightDepth - 1;
                if (len >= (int)lenC.size()) lenC.resize(len + 1, 0.0f);
                lenC[len] += c.lsScore;
                perThread[t].push_back(c.lsScore);
            }
        }
        delete path;
    };
    const int hw = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    // static round-robin assignment: logical thread t runs on worker t % hw
    for (int w = 0; w < std::min(hw, logicalThreads); w++) {
        pool.emplace_back([&, w]() { for (int t = w; t < logicalThreads; t += hw) work(t); });
    }
    for (auto &th : pool) th.join();

    std::vector<float> scores;
    res.lengthContrib.clear();
    for (int t = 0; t < logicalThreads; t++) {
        for (float s : perThread[t]) scores.push_back(s);
        if (perThreadLen[t].size() > res.lengthContrib.size()) res.lengthContrib.resize(perThreadLen[t].size(), 0.0f);
        for (size_t i = 0; i < perThreadLen[t].size(); i++) res.lengthContrib[i] += perThreadLen[t][i];
    }
    mlt_init_finish(scores, numInitSamples, numChains, res);
}

}  // namespace lmc_host
