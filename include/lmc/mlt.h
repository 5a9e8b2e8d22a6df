// mlt.h -- mirror of the reference's integrator surface (src/mlt.h:17-156, src/mlt.cpp:20-215):
//   MarkovState / SplatSample    what of the reference's structs is visible across the ABI
//   MLTInit(...)                 the seeding phase before the chains (src/mlt.h:41-154)
//   MLT(scene, ...)              the whole render: DirectLighting -> MLTInit -> chains (one lmc_ctx per GPU,
//                                chains sharded by global id) -> one NCCL all-reduce of the film -> MergeBuffer ->
//                                BufferToFilm -> WriteImage, with the progressive dump every reportIntervalSpp
// Everything is a thin inline layer over include/lmc/lmc_abi.h; the chain loop itself (the ParallelFor lambda of
// src/mlt.cpp:60-196) runs inside lmc_run_chains on the device.  Errors throw std::runtime_error (the reference's
// Error(), src/flexception.h:23).
#pragma once
#include <chrono>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "lmc_abi.h"
#include "mutation.h"
#include "parsescene.h"

namespace lmc {

struct SplatSample {          // src/mlt.h:25-28
    Float screenPos[2];
    Float contrib[3];
};

// The reference's MarkovState (src/mlt.h:30-39) holds the whole path; across the ABI only what the host loop needs
// from the INIT states is visible: the large-step score each chain starts from (initStates[i].spContrib.lsScore,
// read by the outlier reset, src/mlt.cpp:152-158).  The running states live in HBM (csrc/core/mutation.h MarkovState).
struct MarkovState {
    bool valid = false;
    Float lsScore = Float(0);     // spContrib.lsScore
};

struct MLTResult {
    std::vector<float> film;              // merged H x W x 3 image, what WriteImage wrote
    std::vector<float> indirect;          // summed splats of all chains (indirectBuffer)
    lmc_stats stats;                      // summed over the devices
    Float normalization = Float(0);       // avgScore of MLTInit
    double elapsed = 0.0;                 // seconds of the chain phase (the reference's Tick(timer), src/mlt.cpp:57,200)
    std::string outputNameHDR;
    int intermediateImages = 0;
};

inline void LmcCheck(int rc) { if (rc != LMC_OK) throw std::runtime_error(lmc_last_error()); }

// MLTInit (src/mlt.h:41-154): numInitSamples bidirectional paths, equal-spaced CDF seeding of numChains states.
// ctx == nullptr runs the host restatement (lmc_mlt_init, logical thread count 32 = the reference's machine);
// otherwise the init paths are generated on ctx's GPU (lmc_mlt_init_device, same results for the same
// logicalThreads).  Returns avgScore.
inline Float MLTInit(const Scene *scene, int64_t numInitSamples, int numChains, std::vector<MarkovState> &initStates,
                     lmc_ctx *ctx = nullptr, int logicalThreads = 0) {
    std::vector<float> ls((size_t)numChains);
    float avgScore = 0.0f;
    if (ctx) LmcCheck(lmc_mlt_init_device(ctx, numInitSamples, numChains, logicalThreads > 0 ? logicalThreads : 65536, &avgScore, ls.data()));
    else LmcCheck(lmc_mlt_init(scene->handle, numInitSamples, numChains, logicalThreads > 0 ? logicalThreads : 32, &avgScore, ls.data()));
    initStates.assign((size_t)numChains, MarkovState());
    for (int i = 0; i < numChains; i++) initStates[(size_t)i].lsScore = ls[(size_t)i];
    return avgScore;
}

// MLT() (src/mlt.cpp:20-215).  `devices`: CUDA device indices (one lmc_ctx each).  Options come from
// scene->options (sent to the flattened scene first).  When writeImage is set the merged film is written to
// "<outputName>_timeuse_<elapsed>s.exr" like src/mlt.cpp:208-210 (tone mapping by `hdrmanip` is not reproduced).
inline MLTResult MLT(const Scene *scene, const std::vector<int> &devices = std::vector<int>(1, 0), bool writeImage = true) {
    const DptOptions &opt = *scene->options;
    LmcCheck(opt.ToScene(scene->handle));
    const int W = scene->pixelWidth, H = scene->pixelHeight;
    const int64_t numPixels = (int64_t)W * H;
    const int64_t totalSamples = (int64_t)opt.spp * numPixels;
    const int64_t numChains = opt.numChains;
    const int64_t numSamplesPerChain = totalSamples / numChains;
    // (the reference gives chainId < numSamplesPerChain % numChains one extra sample, src/mlt.cpp:40,64-65 -- its own
    //  comment-level bug, App. B#7; every chain runs numSamplesPerChain here)
    const int G = (int)devices.size();
    if (G < 1) throw std::runtime_error("MLT: no device");
    std::vector<lmc_ctx *> ctx((size_t)G, nullptr);
    std::vector<int32_t> devs(devices.begin(), devices.end());
    LmcCheck(lmc_create_multi(scene->handle, devs.data(), G, ctx.data()));
    MLTResult res;
    struct Guard { std::vector<lmc_ctx *> &c; ~Guard() { for (lmc_ctx *x : c) if (x) lmc_destroy(x); } } guard{ctx};

    // DirectLighting(scene, directBuffer), src/mlt.cpp:33-34
    std::vector<float> direct((size_t)numPixels * 3, 0.0f);
    if (opt.directSpp > 0) LmcCheck(lmc_direct_lighting(ctx[0], opt.directSpp, direct.data()));
    const Float directWeight = opt.directSpp > 0 ? Float(1) / Float(opt.directSpp) : Float(0);

    std::vector<MarkovState> initStates;
    const Float avgScore = MLTInit(scene, opt.numInitSamples, (int)numChains, initStates, ctx[0]);
    std::printf("Average brightness:%g\n", (double)avgScore);
    res.normalization = avgScore;
    std::vector<float> initLs((size_t)numChains);
    for (int64_t i = 0; i < numChains; i++) initLs[(size_t)i] = initStates[(size_t)i].lsScore;

    const auto t0 = std::chrono::system_clock::now();
    for (int g = 0; g < G; g++) {
        lmc_run_desc d = {};
        d.chain_base = (int32_t)(numChains * g / G);
        d.num_chains = (int32_t)(numChains * (g + 1) / G) - d.chain_base;
        d.total_chains = (int32_t)numChains;
        d.samples_per_chain = numSamplesPerChain;
        d.normalization = avgScore;
        LmcCheck(lmc_chains_begin(ctx[g], &d, initLs.data()));
    }
    // chain phase, in slices when progressive dumps are asked for (reportIntervalSpp, src/mlt.cpp:171-193)
    const int64_t slice = opt.reportIntervalSpp > 0 ? std::max<int64_t>(1, (int64_t)opt.reportIntervalSpp * numPixels / numChains)
                                                    : numSamplesPerChain;
    std::vector<float> part((size_t)numPixels * 3), sum((size_t)numPixels * 3);
    for (int64_t done = 0; done < numSamplesPerChain;) {
        const int64_t n = std::min(slice, numSamplesPerChain - done);
        for (int g = 0; g < G; g++) LmcCheck(lmc_run_chains(ctx[g], n, nullptr, nullptr));      // asynchronous: the GPUs overlap
        done += n;
        if (opt.reportIntervalSpp > 0 && done < numSamplesPerChain) {
            std::fill(sum.begin(), sum.end(), 0.0f);
            for (int g = 0; g < G; g++) {
                LmcCheck(lmc_film_read(ctx[g], part.data()));
                for (size_t k = 0; k < sum.size(); k++) sum[k] += part[k];
            }
            const Float sppDone = Float(double(done) * double(numChains) / double(numPixels));
            std::vector<float> film(sum.size());
            LmcCheck(lmc_merge_buffer(direct.data(), directWeight, sum.data(), sppDone > 0 ? Float(1) / sppDone : Float(0), (int64_t)sum.size(), film.data()));
            if (writeImage) LmcCheck(lmc_write_image("intermediate.exr", W, H, film.data()));
            res.intermediateImages++;
        }
    }
    LmcCheck(lmc_allreduce_film(ctx.data(), G));          // the one collective: sum of the per-GPU films
    res.indirect.resize((size_t)numPixels * 3);
    LmcCheck(lmc_film_read(ctx[0], res.indirect.data())); // synchronises ctx[0]; the all-reduce ordered the others
    for (int g = 1; g < G; g++) LmcCheck(lmc_synchronize(ctx[g]));
    res.elapsed = std::chrono::duration<double>(std::chrono::system_clock::now() - t0).count();
    std::printf("Elapsed time:%g\n", res.elapsed);

    res.stats = lmc_stats();
    for (int g = 0; g < G; g++) {
        lmc_stats s;
        LmcCheck(lmc_get_stats(ctx[g], &s));
        for (int k = 0; k < 4; k++) { res.stats.proposed[k] += s.proposed[k]; res.stats.accepted[k] += s.accepted[k]; }
        res.stats.gradient_evals += s.gradient_evals; res.stats.gradient_nonfinite += s.gradient_nonfinite;
        res.stats.kernel_launches += s.kernel_launches; res.stats.outlier_resets += s.outlier_resets;
        res.stats.cache_queries += s.cache_queries; res.stats.cache_hits += s.cache_hits;
    }
    // MergeBuffer(direct / directSpp, indirect / spp) -> BufferToFilm -> WriteImage, src/mlt.cpp:203-210
    const Float sppRun = Float(double(numSamplesPerChain) * double(numChains) / double(numPixels));
    res.film.resize((size_t)numPixels * 3);
    LmcCheck(lmc_merge_buffer(direct.data(), directWeight, res.indirect.data(), sppRun > 0 ? Float(1) / sppRun : Float(0),
                              (int64_t)res.film.size(), res.film.data()));
    res.outputNameHDR = scene->outputName + "_timeuse_" + std::to_string(res.elapsed) + "s.exr";
    if (writeImage) LmcCheck(lmc_write_image(res.outputNameHDR.c_str(), W, H, res.film.data()));
    std::printf("Done!\n");
    return res;
}

}  // namespace lmc
