// oracle_api.cpp -- CPU ORACLE ("Oracle-T", the bit-level twin).  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (langevin-mcmc_b200/liblmc_b200.so) never does and has no CPU
// fallback.
//
// What it is: the reference's chain loop (src/mlt.cpp:60-196), mutations
// (src/mutation_{small,mala,large}.h), path sampling / perturbation (src/path.cpp:529-1449,
// 1953-2160) and scene model, re-stated on the CPU.  The arithmetic lives in
// langevin-mcmc_b200/csrc/core/*.h, written once against IEEE-754 float ops with a shared
// deterministic math header, and is compiled here for x86-64 with `-ffp-contract=off` (no FMA
// contraction, no fast-math) -- the same statements nvcc compiles with `--fmad=false` for
// sm_100a -- so accept/reject sequences can be compared BIT FOR BIT with the GPU.
// What pins it to the reference: tests/test_ref_parity.py checks this restatement's path
// contribution and gradient against the reference's OWN generated code compiled into
// oracle/_ref/ (src/bin/evaluate_path_bidir[_mala]_<c>_<l>_static[_derv]), and
// tests/golden/ holds vectors produced by that code.  Embree, Eigen, OIIO and libm are not
// importable here: closest-hit tie-breaking, texture filtering and transcendental rounding are
// "parity unpinned"  and defined by this oracle.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../langevin-mcmc_b200/csrc/core/chain.h"
#include "../langevin-mcmc_b200/csrc/host/host_scene.h"
#include "../langevin-mcmc_b200/csrc/host/mlt_init.h"

using namespace lmc;

namespace {
struct OScene { lmc_host::SceneStore store; };
thread_local std::string g_err;

struct HostFilm {
    float *p;
    void add(int pix, int c, float v) { p[3 * pix + c] += v; }
};

template <int MAXD>
void run_chains_t(const Scene &sc, int numChains, int chainBase, int totalChains, long long numSteps,
                  long long numSamplesThisChain, float normalization, const float *initLs, float *film,
                  unsigned char *trace, float *aTrace, int threads, unsigned long long *stats) {
    RunParams rp; rp.normalization = normalization; rp.numChains = totalChains;
    rp.numSamplesThisChain = numSamplesThisChain; rp.initLsScore = initLs;
    const int W = sc.cam.width, H = sc.cam.height;
    if (threads < 1) threads = 1;
    std::vector<std::vector<float>> films(threads);
    std::vector<std::vector<unsigned long long>> tstats(threads, std::vector<unsigned long long>(10, 0ULL));
    std::atomic<int> next(0);
    auto work = [&](int w) {
        films[w].assign((size_t)W * H * 3, 0.0f);
        HostFilm hf; hf.p = films[w].data();
        ChainState<MAXD> *cs = new ChainState<MAXD>();
        uint32_t tab[64];
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= numChains) break;
            const int gid = chainBase + i;
            chain_state_init(*cs, initLs ? initLs[gid] : 0.0f);
            chain_run(sc, rp, gid, *cs, numSteps, tab, 1, hf, trace ? trace + (size_t)i * numSteps : nullptr,
                      aTrace ? aTrace + (size_t)i * numSteps : nullptr, 1);
            for (int k = 0; k < 4; k++) { tstats[w][k] += cs->nPropose[k]; tstats[w][4 + k] += cs->nAccept[k]; }
            tstats[w][8] += cs->gradStats[0]; tstats[w][9] += cs->gradStats[1];
        }
        delete cs;
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++) pool.emplace_back(work, w);
    for (auto &t : pool) t.join();
    if (film) for (int w = 0; w < threads; w++) for (size_t k = 0; k < (size_t)W * H * 3; k++) film[k] += films[w][k];
    if (stats) for (int k = 0; k < 10; k++) { stats[k] = 0; for (int w = 0; w < threads; w++) stats[k] += tstats[w][k]; }
}

template <int MAXD>
void bdpt_t(const Scene &sc, int spp, int minDepth, float *film, int threads) {
    const int W = sc.cam.width, H = sc.cam.height;
    const long long total = (long long)spp * W * H;
    std::vector<std::vector<float>> films(threads);
    auto work = [&](int w) {
        films[w].assign((size_t)W * H * 3, 0.0f);
        HostFilm hf; hf.p = films[w].data();
        uint32_t tab[64]; Rng rng; rng.tab = tab; rng.stride = 1; rng_seed(rng, 1000003ULL + (uint64_t)w);
        Path<MAXD> *path = new Path<MAXD>();
        ContribList<Limits<MAXD>::MAXC> contribs;
        for (long long s = w; s < total; s += threads) {
            contribs.clear(); path_clear(*path);
            generate_path_bidir(sc, minDepth, sc.opt.maxDepth, *path, contribs, rng);
            for (int i = 0; i < contribs.n; i++) splat(hf, W, H, contribs.c[i].screenPos, contribs.c[i].contrib);
        }
        delete path;
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++) pool.emplace_back(work, w);
    for (auto &t : pool) t.join();
    // each sample estimates the whole image: per-pixel value = sum * (numPixels / total) = sum / spp
    const float scale = 1.0f / (float)spp;
    for (int w = 0; w < threads; w++) for (size_t k = 0; k < (size_t)W * H * 3; k++) film[k] += films[w][k] * scale;
}
}  // namespace

#define LMCO_TRY try {
#define LMCO_CATCH } catch (const std::exception &e) { g_err = e.what(); return -1; } return 0;

extern "C" {

const char *lmco_last_error() { return g_err.c_str(); }

void *lmco_scene_load(const char *path) {
    try {
        OScene *s = new OScene();
        const std::string p(path);
        if (p.size() > 5 && p.substr(p.size() - 5) == ".pack") lmc_host::load_scene_pack(p, s->store);
        else lmc_host::load_scene_xml(p, s->store);
        return s;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void lmco_scene_free(void *h) { delete (OScene *)h; }
int lmco_scene_save_pack(void *h, const char *path) { LMCO_TRY lmc_host::save_scene_pack(path, ((OScene *)h)->store); LMCO_CATCH }
int lmco_set_option(void *h, const char *name, double v) { return lmc_host::set_option(((OScene *)h)->store.head.opt, name, v) ? 0 : -1; }
int lmco_get_option(void *h, const char *name, double *v) { return lmc_host::get_option(((OScene *)h)->store.head.opt, name, *v) ? 0 : -1; }
// info: width, height, numTris, numNodes, numLights, numGeoms, spp, numInitSamples
int lmco_scene_info(void *h, int *out) {
    const lmc_host::SceneStore &s = ((OScene *)h)->store;
    out[0] = s.head.cam.width; out[1] = s.head.cam.height; out[2] = s.head.numTris; out[3] = s.head.numNodes;
    out[4] = s.head.numLights; out[5] = s.head.numGeoms; out[6] = s.spp; out[7] = s.numInitSamples;
    return 0;
}

int lmco_mlt_init(void *h, long long numInitSamples, int numChains, int logicalThreads, float *normalization, float *initLs) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    lmc_host::InitResult r;
    if (sc.opt.maxDepth <= 4) lmc_host::mlt_init<4>(sc, numInitSamples, numChains, logicalThreads, r);
    else if (sc.opt.maxDepth <= 8) lmc_host::mlt_init<8>(sc, numInitSamples, numChains, logicalThreads, r);
    else if (sc.opt.maxDepth <= 12) lmc_host::mlt_init<12>(sc, numInitSamples, numChains, logicalThreads, r);
    else throw std::runtime_error("maxdepth > 12 is not supported");
    *normalization = r.normalization;
    if (initLs) memcpy(initLs, r.initLsScore.data(), sizeof(float) * numChains);
    LMCO_CATCH
}

// stats[10]: nPropose[4], nAccept[4], gradEvals, gradNonFinite
int lmco_run_chains(void *h, int numChains, int chainBase, int totalChains, long long numSteps,
                    long long numSamplesThisChain, float normalization, const float *initLs, float *film,
                    unsigned char *trace, float *aTrace, int threads, unsigned long long *stats) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    if (sc.opt.maxDepth <= 4) run_chains_t<4>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else if (sc.opt.maxDepth <= 8) run_chains_t<8>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else if (sc.opt.maxDepth <= 12) run_chains_t<12>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else throw std::runtime_error("maxdepth > 12 is not supported");
    LMCO_CATCH
}

// plain bidirectional path tracing estimate of the image (sanity reference for the MLT film)
int lmco_bdpt(void *h, int spp, int minDepth, float *film, int threads) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    if (sc.opt.maxDepth <= 4) bdpt_t<4>(sc, spp, minDepth, film, threads);
    else if (sc.opt.maxDepth <= 8) bdpt_t<8>(sc, spp, minDepth, film, threads);
    else bdpt_t<12>(sc, spp, minDepth, film, threads);
    LMCO_CATCH
}

// rays: n x 6 (org, dir); out: tid[n], tuv[n x 3].  brute != 0 -> O(N) closest hit
int lmco_intersect(void *h, int n, const float *rays, float tmin, float tmax, int brute, int *tid, float *tuv) {
    const Scene sc = ((OScene *)h)->store.view();
    for (int i = 0; i < n; i++) {
        Ray r; r.org = ld3(rays + 6 * i); r.dir = ld3(rays + 6 * i + 3);
        const Hit hit = brute ? brute_closest(sc, r, tmin, tmax) : bvh_traverse<false>(sc, r, tmin, tmax);
        tid[i] = hit.tid; tuv[3 * i] = hit.t; tuv[3 * i + 1] = hit.u; tuv[3 * i + 2] = hit.v;
    }
    return 0;
}

}  // extern "C"
