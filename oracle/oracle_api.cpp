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
#define LMC_ORACLE_HOOK 1
#include <dlfcn.h>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../langevin-mcmc_b200/csrc/core/chain.h"
#include "../langevin-mcmc_b200/csrc/host/host_scene.h"
#include "../langevin-mcmc_b200/csrc/host/mlt_init.h"
#include "../langevin-mcmc_b200/csrc/host/image_decode.h"

using namespace lmc;

namespace {
// ---- Oracle-R: the reference's own generated reverse-mode gradient (oracle/_ref) -----------------
typedef void (*RefDerv)(const float *, const float *, const float *, const float *, float *, float *);
RefDerv g_refDerv[10][9];
bool g_refLoaded = false;
void ref_gradient(int c, int l, const float *lens, const float *primary, const float *sceneSer, float *vert, int nVert,
                  float *grad, int dim) {
    for (int i = 0; i < dim; i++) grad[i] = 0.0f;
    if (c < 1 || c > 9 || l < 0 || l > 8 || !g_refDerv[c][l]) return;
    // The reference reuses one vertParams buffer, so the shape block of an env-map miss holds stale
    // data of an earlier path (src/path.cpp:2547-2550); an all-zero block makes its reverse sweep
    // return NaN.  Mimic the common case: a valid (the first vertex's) triangle.
    if (l == 0) {
        float *blk = vert + 3 + (c - 2) * 59;
        bool zero = true;
        for (int i = 0; i < 46; i++) if (blk[i] != 0.0f) zero = false;
        if (zero) for (int i = 0; i < 46; i++) blk[i] = vert[3 + i];
    }
    (void)nVert;
    g_refDerv[c][l](lens, primary, sceneSer, vert, grad, nullptr);
}
struct OScene { lmc_host::SceneStore store; };
thread_local std::string g_err;

struct HostFilm {
    float *p;
    void add(int pix, int c, float v) { p[3 * pix + c] += v; }
};

bool g_useRef = false;
bool g_cacheGrid = true;   // global-cache queries through the grid (what the device does); false = the linear scan it must equal
bool g_staged = false;     // run the proposal phase through the staged path functions (core/stages.h)
template <int MAXD>
void run_chains_t(const Scene &sc, int numChains, int chainBase, int totalChains, long long numSteps,
                  long long numSamplesThisChain, float normalization, const float *initLs, float *film,
                  unsigned char *trace, float *aTrace, int threads, unsigned long long *stats) {
    RunParams rp; rp.normalization = normalization; rp.numChains = totalChains;
    rp.numSamplesThisChain = numSamplesThisChain; rp.initLsScore = initLs;
    const int W = sc.cam.width, H = sc.cam.height;
    if (threads < 1) threads = 1;
    std::vector<std::vector<float>> films(threads);
    std::vector<std::vector<unsigned long long>> tstats(threads, std::vector<unsigned long long>(18, 0ULL));
    std::atomic<int> next(0);
    auto work = [&](int w) {
        ref_grad_hook() = (g_useRef && g_refLoaded) ? ref_gradient : nullptr;
        films[w].assign((size_t)W * H * 3, 0.0f);
        HostFilm hf; hf.p = films[w].data();
        ChainState<MAXD> *cs = new ChainState<MAXD>();
        H2mcSide *side = new H2mcSide();
        StagedWork<MAXD> *staged = g_staged ? new StagedWork<MAXD>() : nullptr;
        if (staged) memset(staged, 0, sizeof(*staged));
        uint32_t tab[64];
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= numChains) break;
            const int gid = chainBase + i;
            chain_state_init(*cs, initLs ? initLs[gid] : 0.0f);
            memset(side, 0, sizeof(*side));
            chain_run(sc, rp, gid, *cs, numSteps, tab, 1, hf, trace ? trace + (size_t)i * numSteps : nullptr,
                      aTrace ? aTrace + (size_t)i * numSteps : nullptr, 1, side, staged);
            for (int k = 0; k < 4; k++) { tstats[w][k] += cs->nPropose[k]; tstats[w][4 + k] += cs->nAccept[k]; }
            tstats[w][8] += cs->gradStats[0]; tstats[w][9] += cs->gradStats[1]; tstats[w][10] += (unsigned long long)cs->ch.outlierResets;
        }
        delete cs;
        delete side;
        delete staged;
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++) pool.emplace_back(work, w);
    for (auto &t : pool) t.join();
    if (film) for (int w = 0; w < threads; w++) for (size_t k = 0; k < (size_t)W * H * 3; k++) film[k] += films[w][k];
    if (stats) for (int k = 0; k < 18; k++) { stats[k] = 0; for (int w = 0; w < threads; w++) stats[k] += tstats[w][k]; }
}

// Global cache on (option globalcache = 1): chains interact through the cache, so they advance in LOCKSTEP -- every chain
// does iteration k, then the push requests of that iteration are committed in chain order (cache_commit_host) -- which is
// exactly what the device does between its finish and begin kernels.
template <int MAXD>
void run_chains_cache_t(Scene sc, int numChains, int chainBase, int totalChains, long long numSteps,
                        long long numSamplesThisChain, float normalization, const float *initLs, float *film,
                        unsigned char *trace, float *aTrace, int threads, unsigned long long *stats) {
    RunParams rp; rp.normalization = normalization; rp.numChains = totalChains;
    rp.numSamplesThisChain = numSamplesThisChain; rp.initLsScore = initLs;
    std::vector<float> data(LMC_CACHE_FLOATS, 0.0f);
    int count[LMC_CACHE_SLOTS] = {0}, ready[LMC_CACHE_SLOTS] = {0}, gridReady[LMC_CACHE_SLOTS] = {0};
    std::vector<int> grid((size_t)LMC_CACHE_SLOTS * LMC_CACHE_GRID_INTS, 0);
    sc.gc.data = data.data(); sc.gc.count = count; sc.gc.ready = ready;
    sc.gc.grid = g_cacheGrid ? grid.data() : nullptr; sc.gc.gridReady = gridReady;     // lmco_cache_use_grid(0): linear scan
    const int W = sc.cam.width, H = sc.cam.height;
    if (threads < 1) threads = 1;
    std::vector<ChainState<MAXD> *> cs(numChains);
    for (int i = 0; i < numChains; i++) { cs[i] = new ChainState<MAXD>(); chain_state_init(*cs[i], initLs ? initLs[chainBase + i] : 0.0f); }
    std::vector<std::vector<float>> films(threads);
    for (auto &f : films) f.assign((size_t)W * H * 3, 0.0f);
    for (long long k = 0; k < numSteps; k++) {
        std::atomic<int> next(0);
        auto work = [&](int w) {
            ref_grad_hook() = (g_useRef && g_refLoaded) ? ref_gradient : nullptr;
            HostFilm hf; hf.p = films[w].data();
            H2mcSide *side = new H2mcSide(); memset(side, 0, sizeof(*side));
            uint32_t tab[64];
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= numChains) break;
                chain_run(sc, rp, chainBase + i, *cs[i], 1, tab, 1, hf, trace ? trace + (size_t)i * numSteps + k : nullptr,
                          aTrace ? aTrace + (size_t)i * numSteps + k : nullptr, 1, side, (StagedWork<MAXD> *)nullptr);
            }
            delete side;
        };
        std::vector<std::thread> pool;
        for (int w = 0; w < threads; w++) pool.emplace_back(work, w);
        for (auto &t : pool) t.join();
        cache_commit_host<MAXD>(sc, cs.data(), numChains);
    }
    if (film) for (int w = 0; w < threads; w++) for (size_t p = 0; p < (size_t)W * H * 3; p++) film[p] += films[w][p];
    if (stats) {
        for (int k = 0; k < 13; k++) stats[k] = 0;
        for (int i = 0; i < numChains; i++) {
            for (int k = 0; k < 4; k++) { stats[k] += cs[i]->nPropose[k]; stats[4 + k] += cs[i]->nAccept[k]; }
            stats[8] += cs[i]->gradStats[0]; stats[9] += cs[i]->gradStats[1]; stats[10] += (unsigned long long)cs[i]->ch.outlierResets;
            stats[11] += (unsigned long long)cs[i]->ch.cacheQueries; stats[12] += (unsigned long long)cs[i]->ch.cacheHits;
        }
        for (int s = 0; s < LMC_CACHE_SLOTS; s++) stats[13 + s] = (unsigned long long)count[s];
    }
    for (auto *p : cs) delete p;
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

// sink of the DeferredList probe (lmco_deferred_probe): records which flag belongs to which scripted event
namespace {
struct ProbeSink {
    static const bool kDeferConnections = false;
    std::vector<int *> flags; std::vector<int> event; int cur;
    void emit(const Ray &, float, int, int *flag) { flags.push_back(flag); event.push_back(cur); }
    template <class LS> void emit_connection(const Scene &, int, int, int, const LS *, const SurfaceVertex *, const LS &, const SurfaceVertex &, V2,
                                             SubpathContrib *, int *) {}
};
}

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

// stats[18]: nPropose[4], nAccept[4], gradEvals, gradNonFinite, outlierResets, cacheQueries, cacheHits, cacheCount[5]
int lmco_run_chains(void *h, int numChains, int chainBase, int totalChains, long long numSteps,
                    long long numSamplesThisChain, float normalization, const float *initLs, float *film,
                    unsigned char *trace, float *aTrace, int threads, unsigned long long *stats) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    if (sc.opt.cacheEnabled) {      // lockstep runner; H2MC has no cache (src/mutation_h2mc.h)
        if (sc.opt.h2mc) throw std::runtime_error("globalcache applies to the MALA mutation only");
        if (sc.opt.maxDepth <= 4) run_chains_cache_t<4>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
        else if (sc.opt.maxDepth <= 8) run_chains_cache_t<8>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
        else run_chains_cache_t<12>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
        return 0;
    }
    if (sc.opt.maxDepth <= 4) run_chains_t<4>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else if (sc.opt.maxDepth <= 8) run_chains_t<8>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else if (sc.opt.maxDepth <= 12) run_chains_t<12>(sc, numChains, chainBase, totalChains, numSteps, numSamplesThisChain, normalization, initLs, film, trace, aTrace, threads, stats);
    else throw std::runtime_error("maxdepth > 12 is not supported");
    LMCO_CATCH
}

// Step-by-step view of ONE chain for the independent restatement of the chain-level arithmetic in
// tests/test_chain_logic.py (numpy, written from src/mutation_mala.h:83-278, src/mala.cpp:7-51, src/gaussian.cpp:5-36,
// src/mutation_large.h:70-127, src/mlt.cpp:113-170 -- NOT from csrc/core): the chain advances one iteration per record and
// the record holds what the step read and what it left behind.  maxdepth <= 8.  rec[numSteps][LMCO_DBG_STRIDE]:
//   header  0 kind (0 large, 1 isotropic, 2 MALA, 3 H2MC)  1 accepted  2 a  3 dim(cur)  4 dim(prop)  5 cur.ssScore
//           6 prop.ssScore  7 cur.lsScore  8 prop.lsScore  9 cur.scoreSum  10 prop.scoreSum  11 lastScore (before)
//           12 lastScoreSum (before)  13 cur.gaussianInitialized (before)  14 chain.buffered (before)  15 hasContrib
//           16 cur.valid (before)  17 gradient mode of cur (before)  18 gradient mode of prop  19 cur.logDet  20 prop.logDet
//           21 lastScore (after)  22 lastScoreSum (after)  23 chain.t (before)  24 chain.t (after)  25 chain.buffered (after)
//           26 adjacentReject (before)  27 adjacentReject (after)  28 outlier resets during the step
//   arrays (16 floats each, from 32): offset, grad(prop), grad(cur), v1 before, v2 before, v1 after, v2 after,
//           curr_new_v2 before, prop_new_v2 before, cur.mean, cur.invCov_d, cur.covL_d, prop.mean, prop.invCov_d, prop.covL_d
#define LMCO_DBG_STRIDE (32 + 15 * 16)
int lmco_chain_debug(void *h, int chainId, int totalChains, int numSteps, long long numSamplesThisChain, float normalization,
                     const float *initLs, float *rec) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    if (sc.opt.maxDepth > 8 || sc.opt.cacheEnabled) throw std::runtime_error("lmco_chain_debug: maxdepth <= 8, cache off");
    const int MAXD = 8, DIM = 16;
    RunParams rp; rp.normalization = normalization; rp.numChains = totalChains;
    rp.numSamplesThisChain = numSamplesThisChain; rp.initLsScore = initLs;
    ref_grad_hook() = (g_useRef && g_refLoaded) ? ref_gradient : nullptr;
    std::vector<float> film((size_t)sc.cam.width * sc.cam.height * 3, 0.0f);
    HostFilm hf; hf.p = film.data();
    std::unique_ptr<ChainState<MAXD>> csp(new ChainState<MAXD>());
    std::unique_ptr<H2mcSide> side(new H2mcSide());
    ChainState<MAXD> &cs = *csp;
    chain_state_init(cs, initLs ? initLs[chainId] : 0.0f);
    memset(side.get(), 0, sizeof(H2mcSide));
    uint32_t tab[64];
    for (int k = 0; k < numSteps; k++) {
        float *r = rec + (size_t)k * LMCO_DBG_STRIDE;
        for (int i = 0; i < LMCO_DBG_STRIDE; i++) r[i] = 0.0f;
        float *arr = r + 32;
        const int c0 = cs.curIdx;
        const MarkovState<MAXD> &cur = cs.st[c0], &prop = cs.st[c0 ^ 1];
        r[3] = cur.valid ? (float)path_dimension(cur.path) : 0.0f;
        r[5] = cur.sp.ssScore; r[7] = cur.sp.lsScore; r[9] = cur.scoreSum;
        r[11] = cs.ch.lastScore; r[12] = cs.ch.lastScoreSum; r[13] = (float)cur.gaussianInitialized; r[14] = (float)cs.ch.buffered;
        r[16] = (float)cur.valid; r[17] = cur.valid ? (float)mala_grad_mode(sc, cur) : -1.0f;
        r[23] = (float)cs.ch.t; r[26] = (float)cs.ch.adjacentReject;
        const int resets0 = cs.ch.outlierResets;
        for (int i = 0; i < DIM; i++) {
            arr[3 * 16 + i] = cs.ch.v1[i]; arr[4 * 16 + i] = cs.ch.v2[i];
            arr[7 * 16 + i] = cs.ch.curr_new_v2[i]; arr[8 * 16 + i] = cs.ch.prop_new_v2[i];
        }
        if (cur.valid && !cur.gaussianInitialized && mala_grad_mode(sc, cur) == 2) {
            float g[DIM]; for (int i = 0; i < DIM; i++) g[i] = 0.0f;
            mala_eval_gradient(sc, cur, g, nullptr);
            for (int i = 0; i < DIM; i++) arr[2 * 16 + i] = g[i];
        }
        unsigned char tr = 0; float a = 0.0f;
        chain_run(sc, rp, chainId, cs, 1, tab, 1, hf, &tr, &a, 1, side.get(), (StagedWork<MAXD> *)nullptr);
        r[0] = (float)cs.ss.kind; r[1] = (float)((tr >> 2) & 1); r[2] = a; r[15] = (float)cs.ss.hasContrib;
        r[21] = cs.ch.lastScore; r[22] = cs.ch.lastScoreSum; r[24] = (float)cs.ch.t; r[25] = (float)cs.ch.buffered;
        r[27] = (float)cs.ch.adjacentReject; r[28] = (float)(cs.ch.outlierResets - resets0);
        if (cs.ss.hasContrib || cs.ss.kind == STEP_LARGE || a > 0.0f) {
            r[4] = (float)path_dimension(prop.path); r[6] = prop.sp.ssScore; r[8] = prop.sp.lsScore; r[10] = prop.scoreSum;
            r[18] = (float)mala_grad_mode(sc, prop);
        }
        r[19] = cur.gaussian.logDet; r[20] = prop.gaussian.logDet;
        if (cs.ss.kind == STEP_MALA && cs.ss.hasContrib && mala_grad_mode(sc, prop) == 2) {
            float g[DIM]; for (int i = 0; i < DIM; i++) g[i] = 0.0f;
            mala_eval_gradient(sc, prop, g, nullptr);
            for (int i = 0; i < DIM; i++) arr[1 * 16 + i] = g[i];
        }
        for (int i = 0; i < DIM; i++) {
            arr[0 * 16 + i] = cs.ss.offset[i];
            arr[5 * 16 + i] = cs.ch.v1[i]; arr[6 * 16 + i] = cs.ch.v2[i];
            arr[9 * 16 + i] = cur.gaussian.mean[i]; arr[10 * 16 + i] = cur.gaussian.invCov_d[i]; arr[11 * 16 + i] = cur.gaussian.covL_d[i];
            arr[12 * 16 + i] = prop.gaussian.mean[i]; arr[13 * 16 + i] = prop.gaussian.invCov_d[i]; arr[14 * 16 + i] = prop.gaussian.covL_d[i];
        }
    }
    LMCO_CATCH
}

// Oracle-R switch: load oracle/_ref/libpathref_mala.so and route every MALA gradient of subsequent
// lmco_run_chains calls through the reference's generated reverse-mode code (enable = 0 switches back
// to the twin evaluator).  Returns the number of (c, l) functions resolved, or -1.
int lmco_use_reference_gradient(const char *libPath, int enable) {
    if (!enable) { g_useRef = false; return 0; }
    if (!g_refLoaded) {
        void *dl = dlopen(libPath, RTLD_NOW | RTLD_LOCAL);
        if (!dl) { g_err = dlerror(); return -1; }
        int n = 0;
        for (int c = 1; c <= 9; c++) for (int l = 0; l <= 8; l++) {
            char name[96];
            snprintf(name, sizeof(name), "evaluate_path_bidir_mala_%d_%d_static_derv", c, l);
            g_refDerv[c][l] = (RefDerv)dlsym(dl, name);
            if (g_refDerv[c][l]) n++;
        }
        g_refLoaded = n > 0;
        if (!g_refLoaded) { g_err = "no path functions found"; return -1; }
    }
    g_useRef = true;
    int n = 0;
    for (int c = 1; c <= 9; c++) for (int l = 0; l <= 8; l++) if (g_refDerv[c][l]) n++;
    return n;
}

// DirectLighting(scene, buffer) (src/direct.cpp:4-54): film gets the unweighted sample buffer
int lmco_direct_lighting(void *h, int directSpp, float *film, int threads) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    const int W = sc.cam.width, H = sc.cam.height;
    if (direct_lighting_skipped(sc) || directSpp <= 0) return 0;
    const int nX = (W + LMC_DIRECT_TILE - 1) / LMC_DIRECT_TILE, nY = (H + LMC_DIRECT_TILE - 1) / LMC_DIRECT_TILE;
    std::atomic<int> next(0);
    auto work = [&]() {
        HostFilm hf; hf.p = film;           // a tile only splats into its own pixels: no races
        uint32_t tab[64];
        for (;;) {
            const int tile = next.fetch_add(1);
            if (tile >= nX * nY) break;
            Rng rng; rng.tab = tab; rng.stride = 1;
            rng_seed(rng, (uint64_t)(long long)(tile + sc.opt.seedOffset));
            direct_lighting_tile(sc, tile % nX, tile / nX, directSpp, rng, hf);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < (threads < 1 ? 1 : threads); w++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
    LMCO_CATCH
}

// DeferredList (core/stages.h) against the immediate contribution vector on a scripted event sequence.
// type: 0 push (no visibility query), 1 "if (!occluded) push", 2 "if (!occluded) clear", 3 clear.
// occl[i] is the answer of event i's visibility query.  Outputs: the lsScore tags left in each list.
int lmco_deferred_probe(int nEvents, const int *type, const int *occl, float *outImm, int *nImm, float *outDef, int *nDef) {
    Scene sc; memset(&sc, 0, sizeof(sc));
    Ray ray; ray.org = mk3s(0.0f); ray.dir = mk3(0.0f, 0.0f, 1.0f);
    // immediate
    std::vector<float> imm;
    for (int i = 0; i < nEvents; i++) {
        if (type[i] == 0) imm.push_back((float)(i + 1));
        else if (type[i] == 1) { if (!occl[i]) imm.push_back((float)(i + 1)); }
        else if (type[i] == 2) { if (!occl[i]) imm.clear(); }
        else imm.clear();
    }
    // deferred
    const int CAP = 256;
    if (nEvents > CAP) return -1;
    std::vector<SubpathContrib> c(CAP); std::vector<int> flag(CAP); int n = 0;
    ProbeSink sink; DeferredList<ProbeSink> dl; dl.bind(c.data(), flag.data(), &n, CAP, &sink);
    for (int i = 0; i < nEvents; i++) {
        sink.cur = i;
        SubpathContrib x; memset(&x, 0, sizeof(x)); x.lsScore = (float)(i + 1);
        if (type[i] == 0) dl.push(x);
        else if (type[i] == 1) { if (!dl.occluded(sc, ray, 1.0f)) dl.push(x); }
        else if (type[i] == 2) { if (!dl.occluded(sc, ray, 1.0f)) dl.clear(); }
        else dl.clear();
    }
    for (size_t k = 0; k < sink.flags.size(); k++) *sink.flags[k] = cand_resolve(*sink.flags[k], occl[sink.event[k]] != 0);
    n = deferred_compact(c.data(), flag.data(), n);
    *nImm = (int)imm.size(); for (size_t k = 0; k < imm.size(); k++) outImm[k] = imm[k];
    *nDef = n; for (int k = 0; k < n; k++) outDef[k] = c[k].lsScore;
    return 0;
}

// 1: lmco_run_chains runs every proposal through the staged (wavefront) path functions
int lmco_use_staged(int enable) { g_staged = enable != 0; return 0; }
// dropped DeferredList entries since load (core/stages.h capacity invariant): must stay 0
long lmco_deferred_overflows() { return deferred_overflow_count(); }

// plain bidirectional path tracing estimate of the image (sanity reference for the MLT film)
int lmco_bdpt(void *h, int spp, int minDepth, float *film, int threads) {
    LMCO_TRY
    const Scene sc = ((OScene *)h)->store.view();
    if (sc.opt.maxDepth <= 4) bdpt_t<4>(sc, spp, minDepth, film, threads);
    else if (sc.opt.maxDepth <= 8) bdpt_t<8>(sc, spp, minDepth, film, threads);
    else bdpt_t<12>(sc, spp, minDepth, film, threads);
    LMCO_CATCH
}

// Path recorder for the fine-grained parity tests: draws `numLargeSteps` bidirectional paths
// (GeneratePathBidir), turns every contribution into its (c, l) subpath (ToSubpath), optionally
// applies one small-step perturbation (PerturbPathBidir, sigma = perturbstddev) and serializes
// the result in the reference's buffer layout.  Record = LMCO_REC_HEAD + 25 + vstride floats:
//   [c, l, screenX, screenY, lsScore, ssScore, contribRGB(3), perturbed, nVertFloats, dim, pad(4)]
//   primary[25] (time first), vertParams[vstride].   Returns the number of records written.
#define LMCO_REC_HEAD 16
int lmco_sample_paths(void *h, unsigned long long seed, int numLargeSteps, int perturb, int maxLen, int vstride,
                      int maxRecords, float *out) {
    const Scene sc = ((OScene *)h)->store.view();
    const int MAXD = 8;
    uint32_t tab[64]; Rng rng; rng.tab = tab; rng.stride = 1; rng_seed(rng, seed);
    Path<MAXD> *path = new Path<MAXD>(), *sub = new Path<MAXD>(), *pp = new Path<MAXD>();
    ContribList<Limits<MAXD>::MAXC> contribs;
    const int recSize = LMCO_REC_HEAD + 25 + vstride;
    int n = 0;
    const int minDepth = sc.opt.minDepth > 3 ? sc.opt.minDepth : 3;
    auto emit = [&](const Path<MAXD> &p, const SubpathContrib &c, int perturbed) {
        if (n >= maxRecords) return;
        if (c.camDepth + c.lightDepth - 1 > maxLen) return;
        if (serialized_vert_size(c.camDepth, c.lightDepth) > vstride) return;
        float *r = out + (size_t)n * recSize;
        for (int i = 0; i < recSize; i++) r[i] = 0.0f;
        r[0] = (float)c.camDepth; r[1] = (float)c.lightDepth; r[2] = c.screenPos.x; r[3] = c.screenPos.y;
        r[4] = c.lsScore; r[5] = c.ssScore; r[6] = c.contrib.x; r[7] = c.contrib.y; r[8] = c.contrib.z;
        r[9] = (float)perturbed;
        r[10] = (float)serialize_path(sc, p, r + LMCO_REC_HEAD, r + LMCO_REC_HEAD + 25);
        r[11] = (float)path_dimension(p);
        n++;
    };
    for (int s = 0; s < numLargeSteps && n < maxRecords; s++) {
        contribs.clear(); path_clear(*path);
        generate_path_bidir(sc, minDepth, sc.opt.maxDepth, *path, contribs, rng);
        for (int i = 0; i < contribs.n; i++) {
            path_copy(*sub, *path);
            to_subpath(contribs.c[i].camDepth, contribs.c[i].lightDepth, *sub);
            emit(*sub, contribs.c[i], 0);
            if (perturb) {
                float offset[2 * MAXD];
                NormalDist nd = normal_make(0.0f, sc.opt.perturbStdDev);
                const int dim = path_dimension(*sub);
                for (int k = 0; k < dim; k++) offset[k] = normal_draw(nd, rng);
                ContribList<2> pc; pc.clear();
                path_copy(*pp, *sub);
                perturb_path_bidir(sc, offset, *pp, pc, rng);
                if (pc.n > 0) emit(*pp, pc.c[0], 1);
            }
        }
    }
    delete path; delete sub; delete pp;
    return n;
}

// Host twin of lmc_eval_batch: same templated evaluator the kernel compiles (core/pathgrad.h)
int lmco_eval_batch(void *h, int camDepth, int lightDepth, int n, const float *primary, int primaryStride,
                    const float *vertParams, int vertStride, float *logLum, float *grad) {
    const lmc_host::SceneStore &st = ((OScene *)h)->store;
    float sceneSer[38]; memcpy(sceneSer, st.head.sceneSer, sizeof(sceneSer));
    const int dim = primary_param_size(camDepth, lightDepth) - 1;
    for (int i = 0; i < n; i++) {
        const float *p = primary + (size_t)i * primaryStride, *v = vertParams + (size_t)i * vertStride;
        if (grad) logLum[i] = path_loglum_grad_mode(st.head.opt.adjointCompat, camDepth, lightDepth, sceneSer, p, v, grad + (size_t)i * dim);
        else logLum[i] = path_loglum(camDepth, lightDepth, sceneSer, p, v);
    }
    return 0;
}
// gradient + Hessian; hess n x dim x dim row-major.  adjointcompat = 3 selects the forward-over-reverse evaluator (what
// the H2MC mutation runs, LMC_HESS_REV_CHUNK directions per sweep), anything else the second-order forward one
int lmco_eval_batch_hess(void *h, int camDepth, int lightDepth, int n, const float *primary, int primaryStride,
                         const float *vertParams, int vertStride, float *logLum, float *grad, float *hess) {
    const lmc_host::SceneStore &st = ((OScene *)h)->store;
    float sceneSer[38]; memcpy(sceneSer, st.head.sceneSer, sizeof(sceneSer));
    const int dim = primary_param_size(camDepth, lightDepth) - 1;
    if (dim > LMC_HESS_MAXDIM) return -1;
    for (int i = 0; i < n; i++)
        logLum[i] = (st.head.opt.adjointCompat == 3 ? path_loglum_hess_rev<(LMC_HESS_REV_CHUNK > 0 ? LMC_HESS_REV_CHUNK : 4)> : path_loglum_hess)(camDepth, lightDepth, sceneSer, primary + (size_t)i * primaryStride,
                                     vertParams + (size_t)i * vertStride, grad + (size_t)i * dim, hess + (size_t)i * dim * dim);
    return 0;
}
// Jacobi eigen-solver probe: A n x n symmetric row-major -> V (columns = eigenvectors), w ascending
int lmco_jacobi(int n, const float *A, float *V, float *w) {
    float tmp[LMC_H2MC_DIM * LMC_H2MC_DIM];
    if (n < 1 || n > LMC_H2MC_DIM) return -1;
    memcpy(tmp, A, sizeof(float) * n * n);
    jacobi_eigen(n, tmp, V, w);
    return 0;
}
// native image decoders probe (csrc/host/image_decode.h): returns w, h, is8 and the RGB floats
int lmco_decode_image(const char *path, int *whi, float *rgb, long long cap) {
    LMCO_TRY
    lmc_host::DecodedImage im;
    if (!lmc_host::decode_image_native(path, im)) throw std::runtime_error("no native decoder for this extension");
    whi[0] = im.w; whi[1] = im.h; whi[2] = im.is8;
    if (rgb && (long long)im.rgb.size() <= cap) memcpy(rgb, im.rgb.data(), im.rgb.size() * sizeof(float));
    LMCO_CATCH
}
// global_cache_t::query probe: `entries` = PSS_MAX_SIZE x 3 dim floats (pss, v1, v2) of a READY slot
int lmco_cache_query(int dim, const float *entries, const float *pss, float *v1, float *v2) {
    const int s = cache_slot(dim);
    if (s < 0) return -1;
    std::vector<float> data(LMC_CACHE_FLOATS, 0.0f);
    memcpy(data.data() + cache_slot_offset(s), entries, sizeof(float) * (size_t)LMC_CACHE_MAX_SIZE * 3 * dim);
    int count[LMC_CACHE_SLOTS] = {0}, ready[LMC_CACHE_SLOTS] = {0};
    count[s] = LMC_CACHE_MAX_SIZE; ready[s] = 1;
    Scene sc; memset(&sc, 0, sizeof(sc));
    sc.opt.cacheEnabled = 1; sc.gc.data = data.data(); sc.gc.count = count; sc.gc.ready = ready;
    std::vector<int> grid; int gridReady[LMC_CACHE_SLOTS] = {0};
    sc.gc.gridReady = gridReady;
    if (g_cacheGrid) {
        grid.assign((size_t)LMC_CACHE_SLOTS * LMC_CACHE_GRID_INTS, 0);
        sc.gc.grid = grid.data();
        cache_grid_build_host(sc.gc, s);
    }
    return cache_query(sc, dim, pss, v1, v2) ? 1 : 0;
}
// 1 (default): global-cache queries go through the uniform grid of core/scene.h, as on the device; 0: the linear scan
int lmco_cache_use_grid(int enable) { g_cacheGrid = enable != 0; return 0; }
int lmco_scene_serialized(void *h, float *out38) { memcpy(out38, ((OScene *)h)->store.head.sceneSer, 38 * sizeof(float)); return 0; }

// RNG streams from a fresh RNG(seed) each: raw 32-bit draws, uniform_real_distribution<float>(0,1),
// normal_distribution<float>(0,1) (core/rng.h restatement; golden vectors come from the reference header)
int lmco_rng_stream(unsigned long long seed, int n, unsigned int *raw, float *uni, float *nrm) {
    uint32_t tab[64]; Rng r; r.tab = tab; r.stride = 1;
    if (raw) { rng_seed(r, seed); for (int i = 0; i < n; i++) raw[i] = rng_next(r); }
    if (uni) { rng_seed(r, seed); for (int i = 0; i < n; i++) uni[i] = rng_uniform(r); }
    if (nrm) { rng_seed(r, seed); NormalDist d = normal_make(0.0f, 1.0f); for (int i = 0; i < n; i++) nrm[i] = normal_draw(d, r); }
    return 0;
}
// raw draws with the table in lazy mode (lazy = 1) or materialised (0); at draw `zeroAt` (>= 0) the low
// 32 state bits are cleared first, which forces advance_table() (otherwise a 2^-32 event)
int lmco_rng_stream2(unsigned long long seed, int n, int lazy, int zeroAt, unsigned int *raw) {
    uint32_t tab[64]; Rng r; r.tab = tab; r.stride = 1;
    if (lazy) rng_seed_lazy(r, seed); else rng_seed(r, seed);
    for (int i = 0; i < n; i++) {
        if (i == zeroAt) r.state &= ~0xFFFFFFFFULL;
        raw[i] = rng_next(r);
        if (lazy && (i % 97) == 96) {      // persist / restore round trip as the kernels do between launches
            const uint64_t st = r.state; const uint32_t ep = r.epoch;
            rng_restore_lazy(r, seed, st, ep);
        }
    }
    return (int)r.epoch;
}
// deterministic math header probes: fn 0 sin, 1 cos, 2 exp, 3 log, 4 pow(x,y), 5 atan2(x,y), 6 acos, 7 fastlog, 8 fastpow(x,y)
int lmco_math(int fn, int n, const float *x, const float *y, float *out) {
    for (int i = 0; i < n; i++) {
        switch (fn) {
            case 0: out[i] = dm_sin(x[i]); break;
            case 1: out[i] = dm_cos(x[i]); break;
            case 2: out[i] = dm_exp(x[i]); break;
            case 3: out[i] = dm_log(x[i]); break;
            case 4: out[i] = dm_pow(x[i], y[i]); break;
            case 5: out[i] = dm_atan2(x[i], y[i]); break;
            case 6: out[i] = dm_acos(x[i]); break;
            case 7: out[i] = dm_fastlog(x[i]); break;
            case 8: out[i] = dm_fastpow(x[i], y[i]); break;
            default: return -1;
        }
    }
    return 0;
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
