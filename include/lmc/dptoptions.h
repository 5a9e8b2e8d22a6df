// dptoptions.h -- mirror of the reference's `struct DptOptions` (src/dptoptions.h:7-34): same field names and
// defaults, plus the constants the reference compiles in (src/mala.h:9-13, src/global_cache.h:8-14,
// src/mutation.h:5-8) that this library exposes as run-time options.  ToScene() sends every field to an
// lmc_scene under its <dpt> name (src/parsescene.cpp:535-590); FromScene() reads them back.
#pragma once
#include <string>
#include "lmc_abi.h"

namespace lmc {

typedef float Float;   // src/commondef.h:27-28 (SINGLE_PRECISION)

struct DptOptions {
    std::string integrator = "mcmc";                 // MC or MCMC (only the MCMC chain phase is accelerated)
    bool bidirectional = true;
    int spp = 256;
    int numInitSamples = 300000;
    int minDepth = -1;
    int maxDepth = 8;
    int directSpp = 256;

    bool h2mc = false;                               // Hessian-based H2MC kernel
    Float perturbStdDev = Float(0.01);               // H2MC small step sigma
    Float roughnessThreshold = Float(0.05);          // roughness
    Float largeStepProbability = Float(0.05);        // Large step probability
    Float largeStepProbScale = Float(1.0);           // Scale up large step probability in MALA second phase
    bool mala = false;                               // MALA-based kernel
    Float malaGN = Float(100.0);                     // MALA truncated gradient magnitude
    Float malaStepsize = Float(0.005);               // MALA stepsize
    Float malaStdDev = Float(0.005);                 // MALA shrink prior to prevent noisy gradient issue
    bool sampleFromGlobalCache = false;              // LargeStepCache, src/mutation_large_cache.h (not supported: rejected at load time)

    int numChains = 128;
    int seedOffset = 0;
    int reportIntervalSpp = 0;
    Float discreteStdDev = Float(0.01);
    Float uniformMixingProbability = Float(0.1);
    bool useLightCoordinateSampling = false;         // (not supported: rejected at load / create time)
    bool largeStepMultiplexed = false;               // (not supported: rejected at create time)

    // compile-time constants of the reference, run-time options here (same defaults)
    bool globalCache = false;                        // src/global_cache.h: always on in the reference (timing-dependent fill); here an
                                                     // option with a defined fill order, off = every MALA step evaluates its gradient
    int adjointCompat = 1;                           // 1: the reference's reverse sweep (src/chad.cpp:284-287); 0 / 2: true gradient
    int maxDervDepth = 8;                            // src/main.cpp:46
    int pssMinLength = 2, pssMaxLength = 12;         // PSS_MIN_LENGTH / PSS_MAX_LENGTH, src/global_cache.h:8-9
    int outlierWeakRejectCnt = 10000;                // src/mutation.h:6
    int outlierStrongRejectCnt = 1000;               // src/mutation.h:7
    Float outlierRatioThreshold = Float(30.0);       // src/mutation.h:8
    Float lsRatio = Float(0.1);                      // LS_RATIO, src/mala.h:13

    // returns LMC_OK or the first failing lmc_scene_set_option code
    int ToScene(lmc_scene *scene) const {
        struct KV { const char *k; double v; };
        const KV kv[] = {
            {"bidirectional", bidirectional ? 1.0 : 0.0}, {"mindepth", (double)minDepth}, {"maxdepth", (double)maxDepth},
            {"h2mc", h2mc ? 1.0 : 0.0}, {"perturbstddev", perturbStdDev}, {"roughnessthreshold", roughnessThreshold},
            {"largestepprob", largeStepProbability}, {"largestepscale", largeStepProbScale}, {"mala", mala ? 1.0 : 0.0},
            {"mala-gn", malaGN}, {"mala-stepsize", malaStepsize}, {"malastddev", malaStdDev},
            {"numchains", (double)numChains}, {"seedoffset", (double)seedOffset}, {"discretestddev", discreteStdDev},
            {"uniformmixprob", uniformMixingProbability},
            {"uselightcoordinatesampling", useLightCoordinateSampling ? 1.0 : 0.0},
            {"largestepmultiplexed", largeStepMultiplexed ? 1.0 : 0.0},
            {"globalcache", globalCache ? 1.0 : 0.0},
            {"adjointcompat", (double)adjointCompat}, {"maxdervdepth", (double)maxDervDepth},
            {"pssminlength", (double)pssMinLength}, {"pssmaxlength", (double)pssMaxLength},
            {"outlierweakrejectcnt", (double)outlierWeakRejectCnt}, {"outlierstrongrejectcnt", (double)outlierStrongRejectCnt},
            {"outlierratiothreshold", outlierRatioThreshold}, {"lsratio", lsRatio}};
        for (const KV &e : kv) {
            const int rc = lmc_scene_set_option(scene, e.k, e.v);
            if (rc != LMC_OK) return rc;
        }
        return LMC_OK;
    }

    void FromScene(const lmc_scene *scene) {
        auto get = [scene](const char *k, double d) { double v = d; return lmc_scene_get_option(scene, k, &v) == LMC_OK ? v : d; };
        bidirectional = get("bidirectional", 1) != 0; minDepth = (int)get("mindepth", -1); maxDepth = (int)get("maxdepth", 8);
        h2mc = get("h2mc", 0) != 0; perturbStdDev = (Float)get("perturbstddev", 0.01);
        roughnessThreshold = (Float)get("roughnessthreshold", 0.05);
        largeStepProbability = (Float)get("largestepprob", 0.05); largeStepProbScale = (Float)get("largestepscale", 1.0);
        mala = get("mala", 0) != 0; malaGN = (Float)get("mala-gn", 100); malaStepsize = (Float)get("mala-stepsize", 0.005);
        malaStdDev = (Float)get("malastddev", 0.005); numChains = (int)get("numchains", 128); seedOffset = (int)get("seedoffset", 0);
        discreteStdDev = (Float)get("discretestddev", 0.01); uniformMixingProbability = (Float)get("uniformmixprob", 0.1);
        useLightCoordinateSampling = get("uselightcoordinatesampling", 0) != 0;
        largeStepMultiplexed = get("largestepmultiplexed", 0) != 0;
        globalCache = get("globalcache", 0) != 0;
        adjointCompat = (int)get("adjointcompat", 1); maxDervDepth = (int)get("maxdervdepth", 8);
        pssMinLength = (int)get("pssminlength", 2); pssMaxLength = (int)get("pssmaxlength", 12);
        outlierWeakRejectCnt = (int)get("outlierweakrejectcnt", 10000); outlierStrongRejectCnt = (int)get("outlierstrongrejectcnt", 1000);
        outlierRatioThreshold = (Float)get("outlierratiothreshold", 30); lsRatio = (Float)get("lsratio", 0.1);
        lmc_scene_info info;
        if (lmc_scene_get_info(scene, &info) == LMC_OK) {
            spp = info.spp; directSpp = info.direct_spp; numInitSamples = info.num_init_samples;
            reportIntervalSpp = info.report_interval_spp;
        }
    }
};

}  // namespace lmc
