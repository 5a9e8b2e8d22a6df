// abi_consumer.cpp -- a compiled C++ consumer of the drop-in boundary, written the way a maintainer of the
// reference would call it: ParseScene -> DptOptions overrides -> MLT().  Built and run by
// tests/test_abi_consumer.py (g++ against liblmc_b200.so, no CUDA headers needed on the consumer side).
//
// usage: abi_consumer <scene.xml> <out_prefix> <spp> <numChains> <nDevices> [reportIntervalSpp]
// prints one line "RESULT key=value ..." that the test parses.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include "lmc/bsdf.h"
#include "lmc/mlt.h"
#include "lmc/mutation.h"
#include "lmc/parsescene.h"

int main(int argc, char **argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: %s scene.xml out_prefix spp numChains nDevices [reportIntervalSpp]\n", argv[0]); return 2; }
    try {
        std::unique_ptr<lmc::Scene> scene = lmc::ParseScene(argv[1]);
        scene->outputName = argv[2];
        lmc::DptOptions &opt = *scene->options;          // defaults of the xml's <dpt> block, then overrides
        opt.spp = std::atoi(argv[3]);
        opt.numChains = std::atoi(argv[4]);
        opt.maxDepth = 6;
        opt.numInitSamples = 100000;
        opt.directSpp = 4;
        opt.reportIntervalSpp = argc > 6 ? std::atoi(argv[6]) : 0;
        const int nDev = std::atoi(argv[5]);
        std::vector<int> devices;
        for (int i = 0; i < nDev; i++) devices.push_back(i);
        const lmc::MLTResult r = lmc::MLT(scene.get(), devices, true);
        double sum = 0.0; bool finite = true;
        for (float v : r.film) { sum += v; finite = finite && std::isfinite(v); }
        double indirect = 0.0;
        for (float v : r.indirect) indirect += v;
        uint64_t proposed = 0;
        for (int k = 0; k < 4; k++) proposed += r.stats.proposed[k];
        std::printf("RESULT width=%d height=%d mean=%.9g indirect_sum=%.9g finite=%d proposed=%llu mala=%llu grads=%llu norm=%.9g "
                    "intermediate=%d bsdf_size=%d exr=%s\n",
                    scene->pixelWidth, scene->pixelHeight, sum / double(r.film.size()), indirect, finite ? 1 : 0,
                    (unsigned long long)proposed, (unsigned long long)r.stats.proposed[(int)lmc::MutationType::MALASmall],
                    (unsigned long long)r.stats.gradient_evals, (double)r.normalization, r.intermediateImages,
                    lmc::GetMaxBSDFSerializedSize(), r.outputNameHDR.c_str());
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "abi_consumer: %s\n", e.what());
        return 1;
    }
}
