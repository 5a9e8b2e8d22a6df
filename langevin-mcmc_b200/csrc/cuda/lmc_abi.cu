// lmc_abi.cu -- the C ABI (include/lmc/lmc_abi.h) and the sm_100a kernels behind it.
//
// The chain-loop kernels live in chain_kernels.cuh / trace_kernels.cuh (DESIGN.md "Kernels"); here:
//   k_bvh_probe         closest-hit / any-hit harness                       (src/scene.cpp:106-149)
//   k_eval_batch[_hess] log-luminance + PSS gradient (+ Hessian) of serialized paths   (src/path.h:121-125 ABI)
//   k_direct_lighting   DirectLighting(scene, buffer), one thread per tile  (src/direct.cpp:4-54)
// Host code here is glue only: device memory, launches, error translation.  No CPU fallback.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>

#include "../../../include/lmc/lmc_abi.h"
#include "chain_kernels.cuh"
#include "../host/host_scene.h"
#include "../host/mlt_init.h"
#include "../host/image_io.h"
#include "film_comm.h"

using namespace lmc;
using namespace lmc_cuda;

struct lmc_scene { lmc_host::SceneStore store; };

namespace {
thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(LMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// temporary device buffer released on every exit path (the CK macro returns early on a CUDA error)
template <class T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1)); }
    operator T *() const { return p; }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

__global__ void k_bvh_probe(const __grid_constant__ Scene sc, int n, const float *rays, float tmin, float tmax, int anyHit,
                            int *triId, int *geomPrim, float *tuv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r; r.org = ld3(rays + 6 * i); r.dir = ld3(rays + 6 * i + 3);
    if (anyHit) {
        const Hit h = bvh_traverse<true>(sc, r, tmin, tmax);
        triId[i] = h.tid >= 0 ? 1 : 0;
    } else {
        const Hit h = bvh_traverse<false>(sc, r, tmin, tmax);
        triId[i] = h.tid;
        if (geomPrim) { geomPrim[2 * i] = h.tid >= 0 ? sc.tris[h.tid].geom : -1; geomPrim[2 * i + 1] = h.tid >= 0 ? sc.tris[h.tid].prim : -1; }
        if (tuv) { tuv[3 * i] = h.t; tuv[3 * i + 1] = h.u; tuv[3 * i + 2] = h.v; }
    }
}

// one thread = one serialized path (reference ABI: PathFunc / PathFuncDerv, src/path.h:121-125)
__global__ void k_eval_batch(int camDepth, int lightDepth, int n, const float *sceneSer, const float *primary, int primaryStride,
                             const float *vertParams, int vertStride, float *logLum, float *grad, int dim, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = primary + (size_t)i * primaryStride, *v = vertParams + (size_t)i * vertStride;
    if (grad) logLum[i] = path_loglum_grad_mode(mode, camDepth, lightDepth, sceneSer, p, v, grad + (size_t)i * dim);
    else logLum[i] = path_loglum(camDepth, lightDepth, sceneSer, p, v);
}

__global__ void __launch_bounds__(64) k_eval_batch_hess(int camDepth, int lightDepth, int n, const float *sceneSer, const float *primary,
                                                        int primaryStride, const float *vertParams, int vertStride, float *logLum,
                                                        float *grad, float *hess, int dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    logLum[i] = path_loglum_hess(camDepth, lightDepth, sceneSer, primary + (size_t)i * primaryStride,
                                 vertParams + (size_t)i * vertStride, grad + (size_t)i * dim, hess + (size_t)i * dim * dim);
}

// DirectLighting(scene, buffer) (src/direct.cpp:4-54): one CUDA thread per 16 x 16 tile, the tile's RNG and
// sample order as in the reference, so the buffer is reproducible bit for bit (a tile only splats into its
// own pixels)
__global__ void __launch_bounds__(64) k_direct_lighting(const __grid_constant__ Scene sc, int nXTiles, int nYTiles, int directSpp, float *buffer) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= nXTiles * nYTiles) return;
    uint32_t tab[64];
    Rng rng; rng.tab = tab; rng.stride = 1;
    rng_seed_lazy(rng, (uint64_t)(long long)(tile + sc.opt.seedOffset));
    DevFilm df; df.p = buffer;
    direct_lighting_tile(sc, tile % nXTiles, tile / nXTiles, directSpp, rng, df);
}

template <class T> int upload(const std::vector<T> &v, T **out) {
    *out = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CK(cudaMalloc((void **)out, bytes));
    if (!v.empty()) CK(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return LMC_OK;
}
}  // namespace

struct lmc_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    Scene sc;                       // device pointers inside
    std::vector<void *> allocs;
    int maxdTemplate = 8;
    // chains
    void *states = nullptr; size_t stateBytes = 0;
    lmc_run_desc desc{};
    float *initLs = nullptr; int initLsCap = 0;
    bool begun = false;
    // film
    float *film = nullptr; bool filmOwned = false;
    unsigned long long *statsDev = nullptr;
    WaveLists wl{};
    int *listMem = nullptr;
    H2mcSide *sides = nullptr; int sidesCap = 0;
    int listCap = 0;
    WaveCfg wc{};                   // wavefront queues + large-step workspace
    int fullWavesTuned = 0; long long iterationsSinceBegin = 0;
    void *queueMem = nullptr; int queueCap = 0;
    uint64_t launches = 0;
    double lastMs = 0.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *cacheData = nullptr; int *cacheInts = nullptr;       // global cache (option globalcache): entries; count[5] + ready[5] + gridReady[5]
    std::vector<float> initPart; long long initPartKey[4] = {0, 0, 0, 0}; bool initPartValid = false;   // lmc_mlt_init_device_part
    int *cacheGrid = nullptr;                                   // query grids of the ready slots (scene.h LMC_CACHE_GRID_INTS per slot)
    int *cacheBlockCounts = nullptr; int cacheBlockCap = 0;
    ncclComm_t comm = nullptr;      // film all-reduce (lmc_create_multi / lmc_comm_init_rank); NULL for a lone ctx
};

namespace {
int chains_begin(lmc_ctx *c) {
    const int n = c->desc.num_chains;
    const int d = c->maxdTemplate;
    const size_t bytes = (d == 4 ? chain_state_bytes_4() : (d == 8 ? chain_state_bytes_8() : chain_state_bytes_12())) * (size_t)n;
    if (c->states && c->stateBytes != bytes) { cudaFree(c->states); c->states = nullptr; }
    if (!c->states) { CK(cudaMalloc(&c->states, bytes)); c->stateBytes = bytes; }
    void *st = c->states;
    if (c->listCap < n) {
        if (c->listMem) { cudaFree(c->listMem); c->listMem = nullptr; }
        // per sorted list: keys[n] + list[n] + hist/offsets/cursor[NKEYS] + count; plus the large list
        const size_t listLen = (size_t)n + (size_t)LMC_NKEYS * 256;      // + class-alignment gaps (k_sort_scan)
        c->wl.listLen = (int)listLen;
        const int nkeys[3] = {LMC_NKEYS_SMALL, LMC_NKEYS, LMC_NKEYS};
        size_t total = (size_t)n + 4;
        for (int k = 0; k < 3; k++) total += (size_t)n + listLen + 3 * (size_t)nkeys[k] + 4;
        CK(cudaMalloc((void **)&c->listMem, sizeof(int) * total));
        CK(cudaMemsetAsync(c->listMem, 0, sizeof(int) * total, c->stream));
        int *p = c->listMem;
        SortList *sls[3] = {&c->wl.small_, &c->wl.curGrad, &c->wl.propGrad};
        for (int k = 0; k < 3; k++) {
            SortList &sl = *sls[k];
            sl.nkeys = nkeys[k];
            sl.keys = p; p += n; sl.list = p; p += listLen; sl.hist = p; p += nkeys[k]; sl.offsets = p; p += nkeys[k];
            sl.cursor = p; p += nkeys[k]; sl.count = p; p += 4;
        }
        c->wl.large = p; p += n; c->wl.largeCount = p;
        c->listCap = n;
    }
    if (c->queueCap < n || !c->wc.genWork) {
        if (c->queueMem) { cudaFree(c->queueMem); c->queueMem = nullptr; }
        if (c->wc.genWork) { cudaFree(c->wc.genWork); c->wc.genWork = nullptr; }
        // 2 sets x 2 step kinds: a buffer of n entries (chain, payload, hit) shared by the light- and the
        // camera-subpath queue of that kind; a shadow queue of 4n segments; 9 counters
        const size_t nn = (size_t)n, shCap = 4 * nn;
        const size_t perBuf = nn * (sizeof(int) + sizeof(PayloadLite) + sizeof(float4)) + 64;
        const size_t cqCap = 4 * nn;
        const size_t bytes = 4 * perBuf + shCap * (2 * sizeof(float4) + sizeof(int *) + sizeof(int)) + cqCap * sizeof(int4) + 1024;
        CK(cudaMalloc(&c->queueMem, bytes));
        char *p = (char *)c->queueMem;
        auto take = [&](size_t b) { char *r = p; p += (b + 15) & ~(size_t)15; return r; };
        for (int s = 0; s < 2; s++) for (int kind = 0; kind < 2; kind++) {
            uint4 *payload = (uint4 *)take(nn * sizeof(PayloadLite));
            float4 *hit = (float4 *)take(nn * sizeof(float4));
            int *chain = (int *)take(nn * sizeof(int));
            for (int end = 0; end < 2; end++) {       // light-subpath queue grows up, camera-subpath queue grows down
                RayQueue &q = c->wc.wq.q[s][2 * kind + end];
                q.payload = payload; q.hit = hit; q.chain = chain; q.cap = n;
                q.base = end ? n - 1 : 0; q.dirn = end ? -1 : 1;
            }
        }
        c->wc.wq.sh.org = (float4 *)take(shCap * sizeof(float4)); c->wc.wq.sh.dir = (float4 *)take(shCap * sizeof(float4));
        c->wc.wq.sh.flag = (int **)take(shCap * sizeof(int *));
        c->wc.wq.sh.chain = (int *)take(shCap * sizeof(int));
        c->wc.wq.cq.item = (int4 *)take(cqCap * sizeof(int4));
        c->wc.wq.cq.cap = (int)cqCap;
        c->wc.queueCounts = (int *)take(LMC_NCOUNTERS * sizeof(int));
        for (int s = 0; s < 2; s++) for (int k = 0; k < 4; k++) c->wc.wq.q[s][k].count = c->wc.queueCounts + 4 * s + k;
        c->wc.wq.sh.count = c->wc.queueCounts + 8;
        c->wc.wq.cq.count = c->wc.queueCounts + 12;
        c->wc.wq.sh.cap = (int)shCap;
        CK(cudaMemsetAsync(c->wc.queueCounts, 0, LMC_NCOUNTERS * sizeof(int), c->stream));
        const size_t gwBytes = (d == 4 ? gen_work_bytes_4() : (d == 8 ? gen_work_bytes_8() : gen_work_bytes_12())) * nn;
        CK(cudaMalloc(&c->wc.genWork, gwBytes));
        c->queueCap = n;
    }
    if (c->sc.opt.h2mc) {
        if (c->sidesCap < n) {
            if (c->sides) { cudaFree(c->sides); c->sides = nullptr; }
            CK(cudaMalloc((void **)&c->sides, sizeof(H2mcSide) * (size_t)n));
            c->sidesCap = n;
        }
        CK(cudaMemsetAsync(c->sides, 0, sizeof(H2mcSide) * (size_t)n, c->stream));
        if (!c->wc.padSide) { CK(cudaMalloc((void **)&c->wc.padSide, sizeof(H2mcSide))); CK(cudaMemsetAsync(c->wc.padSide, 0, sizeof(H2mcSide), c->stream)); }
    }
    if (c->sc.opt.cacheEnabled) {
        // a fresh cache for every job (GlobalCache globalCache; src/mlt.cpp:53)
        CK(cudaMemsetAsync(c->cacheData, 0, sizeof(float) * (size_t)LMC_CACHE_FLOATS, c->stream));
        CK(cudaMemsetAsync(c->cacheInts, 0, sizeof(int) * 3 * LMC_CACHE_SLOTS, c->stream));
        const int need = LMC_CACHE_SLOTS * ((n + LMC_CACHE_BLOCK - 1) / LMC_CACHE_BLOCK) + 1;
        if (c->cacheBlockCap < need) {
            if (c->cacheBlockCounts) { cudaFree(c->cacheBlockCounts); c->cacheBlockCounts = nullptr; }
            CK(cudaMalloc((void **)&c->cacheBlockCounts, sizeof(int) * (size_t)need));
            c->cacheBlockCap = need;
        }
        c->wc.cacheBlockCounts = c->cacheBlockCounts;
    }
    c->fullWavesTuned = 0; c->iterationsSinceBegin = 0;
    c->wc.fullWavesTuned = &c->fullWavesTuned; c->wc.iterationsSinceBegin = &c->iterationsSinceBegin;
    c->launches++;
    CK(d == 4 ? launch_chain_init_4(c->stream, st, n, c->desc.chain_base, c->initLs)
              : (d == 8 ? launch_chain_init_8(c->stream, st, n, c->desc.chain_base, c->initLs)
                        : launch_chain_init_12(c->stream, st, n, c->desc.chain_base, c->initLs)));
    return LMC_OK;
}

int run_chains(lmc_ctx *c, long long numSteps, unsigned char *dTrace, float *dATrace) {
    const int n = c->desc.num_chains;
    const int d = c->maxdTemplate;
    RunParams rp; rp.normalization = c->desc.normalization; rp.numChains = c->desc.total_chains;
    rp.numSamplesThisChain = c->desc.samples_per_chain; rp.initLsScore = c->initLs;
    void *st = c->states;
    CK(cudaEventRecord(c->ev0, c->stream));
    unsigned long long nl = 0;
    CK(d == 4 ? launch_chain_run_4(c->stream, c->sc, rp, c->desc.chain_base, st, n, numSteps, c->film, dTrace, dATrace, c->wl, c->wc, &nl, c->sides)
              : (d == 8 ? launch_chain_run_8(c->stream, c->sc, rp, c->desc.chain_base, st, n, numSteps, c->film, dTrace, dATrace, c->wl, c->wc, &nl, c->sides)
                        : launch_chain_run_12(c->stream, c->sc, rp, c->desc.chain_base, st, n, numSteps, c->film, dTrace, dATrace, c->wl, c->wc, &nl, c->sides)));
    c->launches += nl;
    CK(cudaEventRecord(c->ev1, c->stream));
    return LMC_OK;
}

int chain_stats(lmc_ctx *c) {
    const int n = c->desc.num_chains;
    const int d = c->maxdTemplate;
    const void *st = c->states;
    CK(cudaMemsetAsync(c->statsDev, 0, 13 * sizeof(unsigned long long), c->stream));
    c->launches++;
    CK(d == 4 ? launch_chain_stats_4(c->stream, st, n, c->statsDev)
              : (d == 8 ? launch_chain_stats_8(c->stream, st, n, c->statsDev) : launch_chain_stats_12(c->stream, st, n, c->statsDev)));
    return LMC_OK;
}
}  // namespace

extern "C" {

const char *lmc_last_error(void) { return g_err.c_str(); }
const char *lmc_version(void) { return "lmc-b200 0.1 (sm_100a)"; }

int lmc_scene_load(const char *path, lmc_scene **out) {
    if (!path || !out) return fail(LMC_ERR_ARG, "null argument");
    try {
        lmc_scene *s = new lmc_scene();
        const std::string p(path);
        if (p.size() > 5 && p.substr(p.size() - 5) == ".pack") lmc_host::load_scene_pack(p, s->store);
        else lmc_host::load_scene_xml(p, s->store);
        *out = s;
    } catch (const std::exception &e) { return fail(LMC_ERR_IO, e.what()); }
    return LMC_OK;
}
int lmc_scene_save_pack(const lmc_scene *scene, const char *path) {
    if (!scene || !path) return fail(LMC_ERR_ARG, "null argument");
    try { lmc_host::save_scene_pack(path, scene->store); } catch (const std::exception &e) { return fail(LMC_ERR_IO, e.what()); }
    return LMC_OK;
}
void lmc_scene_free(lmc_scene *scene) { delete scene; }
int lmc_scene_get_info(const lmc_scene *scene, lmc_scene_info *out) {
    if (!scene || !out) return fail(LMC_ERR_ARG, "null argument");
    const lmc_host::SceneStore &s = scene->store;
    out->width = s.head.cam.width; out->height = s.head.cam.height; out->num_triangles = s.head.numTris;
    out->num_bvh_nodes = s.head.numNodes; out->num_lights = s.head.numLights; out->num_shapes = s.head.numGeoms;
    out->num_textures = s.head.numTextures; out->spp = s.spp; out->direct_spp = s.directSpp; out->num_init_samples = s.numInitSamples;
    out->report_interval_spp = s.reportIntervalSpp;
    return LMC_OK;
}
int lmc_scene_set_option(lmc_scene *scene, const char *name, double value) {
    if (!scene || !name) return fail(LMC_ERR_ARG, "null argument");
    if (!lmc_host::set_option(scene->store.head.opt, name, value)) return fail(LMC_ERR_ARG, std::string("Unknown dpt option:") + name);
    return LMC_OK;
}
int lmc_scene_get_option(const lmc_scene *scene, const char *name, double *value) {
    if (!scene || !name || !value) return fail(LMC_ERR_ARG, "null argument");
    if (!lmc_host::get_option(scene->store.head.opt, name, *value)) return fail(LMC_ERR_ARG, std::string("Unknown dpt option:") + name);
    return LMC_OK;
}
int lmc_scene_serialized(const lmc_scene *scene, float *out38) {
    if (!scene || !out38) return fail(LMC_ERR_ARG, "null argument");
    memcpy(out38, scene->store.head.sceneSer, 38 * sizeof(float));
    out38[0] = scene->store.head.opt.useLightCoordinateSampling ? 1.0f : 0.0f;
    return LMC_OK;
}

int lmc_mlt_init(const lmc_scene *scene, int64_t num_init_samples, int32_t num_chains, int32_t logical_threads,
                 float *normalization, float *init_ls_score) {
    if (!scene || !normalization || num_chains <= 0 || num_init_samples <= 0) return fail(LMC_ERR_ARG, "bad argument");
    try {
        const Scene sc = scene->store.view();
        lmc_host::InitResult r;
        if (sc.opt.maxDepth <= 4) lmc_host::mlt_init<4>(sc, num_init_samples, num_chains, logical_threads, r);
        else if (sc.opt.maxDepth <= 8) lmc_host::mlt_init<8>(sc, num_init_samples, num_chains, logical_threads, r);
        else if (sc.opt.maxDepth <= 12) lmc_host::mlt_init<12>(sc, num_init_samples, num_chains, logical_threads, r);
        else return fail(LMC_ERR_UNSUPPORTED, "maxdepth > 12 is not supported");
        *normalization = r.normalization;
        if (init_ls_score) memcpy(init_ls_score, r.initLsScore.data(), sizeof(float) * (size_t)num_chains);
    } catch (const std::exception &e) { return fail(LMC_ERR_STATE, e.what()); }
    return LMC_OK;
}

// Init paths of the logical threads [t0, t1) on this ctx's GPU: lsScores of their contributions in (thread, sample,
// contribution) order.  Two passes of k_mlt_init_paths (count, then emit at the scanned offsets).
static int mlt_init_device_scores(lmc_ctx *c, int64_t num_init_samples, int32_t logical_threads, int32_t t0, int32_t t1,
                                  std::vector<float> &scores) {
    CK(cudaSetDevice(c->device));
    const int d = c->maxdTemplate;
    const int T = t1 - t0;
    scores.clear();
    if (T <= 0) return LMC_OK;
    DevBuf<int> dCounts; DevBuf<long long> dOffsets; DevBuf<float> dScores;
    CK(dCounts.alloc((size_t)T));
    CK(dOffsets.alloc((size_t)T));
    auto launch = [&](int emit) {
        return d == 4 ? launch_mlt_init_paths_4(c->stream, c->sc, num_init_samples, logical_threads, t0, t1, emit, dCounts, dOffsets, dScores)
                      : (d == 8 ? launch_mlt_init_paths_8(c->stream, c->sc, num_init_samples, logical_threads, t0, t1, emit, dCounts, dOffsets, dScores)
                                : launch_mlt_init_paths_12(c->stream, c->sc, num_init_samples, logical_threads, t0, t1, emit, dCounts, dOffsets, dScores));
    };
    std::vector<int> counts(T);
    std::vector<long long> offsets(T);
    cudaError_t e = launch(0);
    c->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), dCounts, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    long long total = 0;
    if (e == cudaSuccess) {
        for (int t = 0; t < T; t++) { offsets[t] = total; total += counts[t]; }
        scores.resize((size_t)total);
        e = dScores.alloc((size_t)total);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dOffsets, offsets.data(), sizeof(long long) * (size_t)T, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) { e = launch(1); c->launches++; }
    if (e == cudaSuccess && total > 0) e = cudaMemcpyAsync(scores.data(), dScores, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(LMC_ERR_CUDA, std::string("lmc_mlt_init_device: ") + cudaGetErrorString(e));
    return LMC_OK;
}

int lmc_mlt_init_device(lmc_ctx *c, int64_t num_init_samples, int32_t num_chains, int32_t logical_threads,
                        float *normalization, float *init_ls_score) {
    if (!c || !normalization || num_chains <= 0 || num_init_samples <= 0 || logical_threads <= 0) return fail(LMC_ERR_ARG, "bad argument");
    std::vector<float> scores;
    const int rc = mlt_init_device_scores(c, num_init_samples, logical_threads, 0, logical_threads, scores);
    if (rc) return rc;
    return lmc_mlt_init_finish(scores.data(), (int64_t)scores.size(), num_init_samples, num_chains, normalization, init_ls_score);
}

int lmc_mlt_init_device_part(lmc_ctx *c, int64_t num_init_samples, int32_t logical_threads, int32_t thread_begin, int32_t thread_end,
                             float *scores, int64_t capacity, int64_t *num_scores) {
    if (!c || !num_scores || num_init_samples <= 0 || logical_threads <= 0 || thread_begin < 0 || thread_end > logical_threads ||
        thread_begin > thread_end) return fail(LMC_ERR_ARG, "bad argument");
    // the part is generated on the first call (scores == NULL asks for its size) and kept until it has been copied out
    if (!(c->initPartValid && c->initPartKey[0] == num_init_samples && c->initPartKey[1] == logical_threads &&
          c->initPartKey[2] == thread_begin && c->initPartKey[3] == thread_end)) {
        const int rc = mlt_init_device_scores(c, num_init_samples, logical_threads, thread_begin, thread_end, c->initPart);
        if (rc) return rc;
        c->initPartKey[0] = num_init_samples; c->initPartKey[1] = logical_threads; c->initPartKey[2] = thread_begin; c->initPartKey[3] = thread_end;
        c->initPartValid = true;
    }
    *num_scores = (int64_t)c->initPart.size();
    if (!scores) return LMC_OK;
    if (capacity < *num_scores) return fail(LMC_ERR_ARG, "lmc_mlt_init_device_part: capacity too small");
    if (!c->initPart.empty()) memcpy(scores, c->initPart.data(), sizeof(float) * c->initPart.size());
    c->initPart.clear(); c->initPart.shrink_to_fit(); c->initPartValid = false;
    return LMC_OK;
}

int lmc_mlt_init_finish(const float *scores, int64_t num_scores, int64_t num_init_samples, int32_t num_chains,
                        float *normalization, float *init_ls_score) {
    if ((!scores && num_scores > 0) || num_scores < 0 || !normalization || num_chains <= 0 || num_init_samples <= 0) return fail(LMC_ERR_ARG, "bad argument");
    try {
        const std::vector<float> v(scores, scores + num_scores);
        lmc_host::InitResult r;
        lmc_host::mlt_init_finish(v, num_init_samples, num_chains, r);
        *normalization = r.normalization;
        if (init_ls_score) memcpy(init_ls_score, r.initLsScore.data(), sizeof(float) * (size_t)num_chains);
    } catch (const std::exception &ex) { return fail(LMC_ERR_STATE, ex.what()); }
    return LMC_OK;
}

int lmc_direct_lighting(lmc_ctx *c, int32_t direct_spp, float *host_rgb) {
    if (!c || !host_rgb || direct_spp < 0) return fail(LMC_ERR_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    const int W = c->sc.cam.width, H = c->sc.cam.height;
    const size_t bytes = (size_t)W * H * 3 * sizeof(float);
    if (direct_lighting_skipped(c->sc) || direct_spp == 0) { memset(host_rgb, 0, bytes); return LMC_OK; }
    DevBuf<float> dBuf;
    CK(dBuf.alloc(bytes / sizeof(float)));
    cudaError_t e = cudaMemsetAsync(dBuf, 0, bytes, c->stream);
    const int nX = (W + LMC_DIRECT_TILE - 1) / LMC_DIRECT_TILE, nY = (H + LMC_DIRECT_TILE - 1) / LMC_DIRECT_TILE;
    if (e == cudaSuccess) {
        k_direct_lighting<<<(nX * nY + 63) / 64, 64, 0, c->stream>>>(c->sc, nX, nY, direct_spp, dBuf);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_rgb, dBuf, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(LMC_ERR_CUDA, std::string("lmc_direct_lighting: ") + cudaGetErrorString(e));
    return LMC_OK;
}

int lmc_create(const lmc_scene *scene, int32_t device, lmc_ctx **out) {
    if (!scene || !out) return fail(LMC_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(LMC_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(LMC_ERR_ARG, "device index out of range");
    CK(cudaSetDevice(device));
    const lmc_host::SceneStore &s = scene->store;
    if (s.head.opt.maxDepth < 2 || s.head.opt.maxDepth > 12) return fail(LMC_ERR_UNSUPPORTED, "maxdepth must be in [2, 12]");
    if (s.head.opt.h2mc && s.head.opt.maxDepth > 8) return fail(LMC_ERR_UNSUPPORTED, "h2mc needs maxdepth <= 8 (dense Gaussians are stored for dim <= 16)");
    if (s.head.opt.maxDervDepth > 8) return fail(LMC_ERR_UNSUPPORTED, "maxdervdepth must be <= 8 (derivative functions exist for camDepth + lightDepth - 1 <= 8)");
    if (s.head.opt.largeStepMultiplexed) return fail(LMC_ERR_UNSUPPORTED, "largestepmultiplexed is not supported");
    if (s.head.opt.useLightCoordinateSampling) return fail(LMC_ERR_UNSUPPORTED, "uselightcoordinatesampling is not supported (the sampler has no light-coordinate branch)");
    if (s.head.opt.cacheEnabled && s.head.opt.h2mc) return fail(LMC_ERR_UNSUPPORTED, "globalcache applies to the MALA mutation only");
    if (s.head.opt.adjointCompat < 0 || s.head.opt.adjointCompat > 2) return fail(LMC_ERR_ARG, "adjointcompat must be 0, 1 or 2");
    lmc_ctx *c = new lmc_ctx();
    c->device = device;
    c->sc = s.head;
    c->maxdTemplate = s.head.opt.maxDepth <= 4 ? 4 : (s.head.opt.maxDepth <= 8 ? 8 : 12);
    Scene &d = c->sc;
    int rc = LMC_OK;
#define UP(field, vec, T) do { T *p_ = nullptr; rc = upload<T>(vec, &p_); if (rc) { lmc_destroy(c); return rc; } c->allocs.push_back(p_); field = p_; } while (0)
    UP(d.tris, s.tris, TriGeom); UP(d.shade, s.shade, TriShade); UP(d.nodes, s.nodes, BvhNode); UP(d.mats, s.mats, Material);
    UP(d.textures, s.textures, Texture); UP(d.texData, s.texData, float); UP(d.lights, s.lights, Light);
    UP(d.lightPickCdf, s.lightPickCdf, float); UP(d.lightCdf, s.lightCdf, float); UP(d.lightPrimTid, s.lightPrimTid, int);
    UP(d.env.image, s.envImage, float); UP(d.env.cdfRows, s.envCdfRows, float); UP(d.env.cdfCols, s.envCdfCols, float);
    UP(d.env.rowWeights, s.envRowWeights, float);
#undef UP
    d.gc.data = nullptr; d.gc.count = nullptr; d.gc.ready = nullptr; d.gc.grid = nullptr; d.gc.gridReady = nullptr;
    if (d.opt.cacheEnabled) {
        if (cudaMalloc((void **)&c->cacheData, sizeof(float) * (size_t)LMC_CACHE_FLOATS) != cudaSuccess ||
            cudaMalloc((void **)&c->cacheInts, sizeof(int) * 3 * LMC_CACHE_SLOTS) != cudaSuccess ||
            cudaMalloc((void **)&c->cacheGrid, sizeof(int) * (size_t)LMC_CACHE_SLOTS * LMC_CACHE_GRID_INTS) != cudaSuccess) {
            lmc_destroy(c);
            return fail(LMC_ERR_CUDA, "device allocation failed (global cache)");
        }
        d.gc.data = c->cacheData; d.gc.count = c->cacheInts; d.gc.ready = c->cacheInts + LMC_CACHE_SLOTS;
        d.gc.gridReady = c->cacheInts + 2 * LMC_CACHE_SLOTS; d.gc.grid = c->cacheGrid;
    }
    const size_t filmBytes = (size_t)d.cam.width * d.cam.height * 3 * sizeof(float);
    if (cudaMalloc((void **)&c->film, filmBytes) != cudaSuccess || cudaMemset(c->film, 0, filmBytes) != cudaSuccess ||
        cudaMalloc((void **)&c->statsDev, 13 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        lmc_destroy(c);
        return fail(LMC_ERR_CUDA, "device allocation failed");
    }
    c->filmOwned = true;
    {
        // Both device forms of the proposal phase give identical chains; which one is faster depends on the
        // chain count (tools/smalln.sh: the per-vertex wavefront wins from ~4e5 chains up, below that its ~50
        // small launches per iteration cost more than its SIMD efficiency gains).  LMC_WAVEFRONT=0/1 forces one.
        if (!getenv("LMC_NO_AUX_STREAM") &&
            (cudaStreamCreateWithFlags(&c->wc.aux, cudaStreamNonBlocking) != cudaSuccess ||
             cudaEventCreateWithFlags(&c->wc.evFork, cudaEventDisableTiming) != cudaSuccess ||
             cudaEventCreateWithFlags(&c->wc.evJoin, cudaEventDisableTiming) != cudaSuccess)) {
            lmc_destroy(c);
            return fail(LMC_ERR_CUDA, "stream / event creation failed");
        }
        const char *wf = getenv("LMC_WAVEFRONT");
        c->wc.wavefront = wf ? ((wf[0] == '0') ? 0 : 1) : -1;
        cudaDeviceProp prop;
        c->wc.smCount = (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ? prop.multiProcessorCount : 148;
    }
    *out = c;
    return LMC_OK;
}

void lmc_destroy(lmc_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (void *p : c->allocs) cudaFree(p);
    if (c->states) cudaFree(c->states);
    if (c->initLs) cudaFree(c->initLs);
    if (c->film && c->filmOwned) cudaFree(c->film);
    if (c->statsDev) cudaFree(c->statsDev);
    if (c->cacheData) cudaFree(c->cacheData);
    if (c->cacheInts) cudaFree(c->cacheInts);
    if (c->cacheGrid) cudaFree(c->cacheGrid);
    if (c->cacheBlockCounts) cudaFree(c->cacheBlockCounts);
    if (c->listMem) cudaFree(c->listMem);
    if (c->sides) cudaFree(c->sides);
    if (c->queueMem) cudaFree(c->queueMem);
    if (c->wc.genWork) cudaFree(c->wc.genWork);
    if (c->wc.padSide) cudaFree(c->wc.padSide);
    if (c->wc.aux) cudaStreamDestroy(c->wc.aux);
    if (c->wc.evFork) cudaEventDestroy(c->wc.evFork);
    if (c->wc.evJoin) cudaEventDestroy(c->wc.evJoin);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->comm && nccl_api().lib) nccl_api().CommDestroy(c->comm);
    delete c;
}

int lmc_set_stream(lmc_ctx *c, void *cuda_stream) {
    if (!c) return fail(LMC_ERR_ARG, "null ctx");
    c->stream = (cudaStream_t)cuda_stream;
    return LMC_OK;
}

int lmc_chains_begin(lmc_ctx *c, const lmc_run_desc *desc, const float *init_ls_score) {
    if (!c || !desc) return fail(LMC_ERR_ARG, "null argument");
    if (desc->num_chains <= 0 || desc->total_chains < desc->num_chains || desc->chain_base < 0 ||
        desc->chain_base + desc->num_chains > desc->total_chains) return fail(LMC_ERR_ARG, "inconsistent run descriptor");
    CK(cudaSetDevice(c->device));
    c->desc = *desc;
    if (c->initLs && c->initLsCap < desc->total_chains) { cudaFree(c->initLs); c->initLs = nullptr; }
    if (!c->initLs) { CK(cudaMalloc((void **)&c->initLs, sizeof(float) * (size_t)desc->total_chains)); c->initLsCap = desc->total_chains; }
    if (init_ls_score) CK(cudaMemcpyAsync(c->initLs, init_ls_score, sizeof(float) * (size_t)desc->total_chains, cudaMemcpyHostToDevice, c->stream));
    else CK(cudaMemsetAsync(c->initLs, 0, sizeof(float) * (size_t)desc->total_chains, c->stream));
    const int rc = chains_begin(c);
    if (rc) return rc;
    c->begun = true;
    return lmc_film_clear(c);
}

int lmc_run_chains(lmc_ctx *c, int64_t num_mutations, uint8_t *trace, float *a_trace) {
    if (!c) return fail(LMC_ERR_ARG, "null ctx");
    if (!c->begun) return fail(LMC_ERR_STATE, "lmc_chains_begin has not been called");
    if (num_mutations <= 0) return fail(LMC_ERR_ARG, "num_mutations must be positive");
    CK(cudaSetDevice(c->device));
    DevBuf<unsigned char> dTrace; DevBuf<float> dA;
    const size_t cnt = (size_t)c->desc.num_chains * (size_t)num_mutations;
    if (trace) CK(dTrace.alloc(cnt));
    if (a_trace) CK(dA.alloc(cnt));
    int rc = run_chains(c, num_mutations, dTrace, dA);
    if (rc == LMC_OK && (trace || a_trace)) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess && trace) e = cudaMemcpy(trace, dTrace, cnt, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && a_trace) e = cudaMemcpy(a_trace, dA, cnt * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(LMC_ERR_CUDA, cudaGetErrorString(e));
    }
    return rc;
}

int lmc_synchronize(lmc_ctx *c) {
    if (!c) return fail(LMC_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return LMC_OK;
}

int lmc_get_stats(lmc_ctx *c, lmc_stats *out) {
    if (!c || !out) return fail(LMC_ERR_ARG, "null argument");
    memset(out, 0, sizeof(*out));
    CK(cudaSetDevice(c->device));
    if (c->begun) {
        const int rc = chain_stats(c);
        if (rc) return rc;
        unsigned long long h[13];
        CK(cudaMemcpyAsync(h, c->statsDev, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 4; k++) { out->proposed[k] = h[k]; out->accepted[k] = h[4 + k]; }
        out->gradient_evals = h[8]; out->gradient_nonfinite = h[9]; out->outlier_resets = h[10];
        out->cache_queries = h[11]; out->cache_hits = h[12];
        if (c->sc.opt.cacheEnabled) {
            int ci[2 * LMC_CACHE_SLOTS];
            CK(cudaMemcpyAsync(ci, c->cacheInts, sizeof(ci), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            for (int s = 0; s < LMC_CACHE_SLOTS; s++) out->cache_count[s] = (uint32_t)ci[s];
        }
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->lastMs = ms;
        else cudaGetLastError();
    }
    out->kernel_launches = c->launches;
    out->last_kernel_ms = c->lastMs;
    return LMC_OK;
}

int lmc_film_clear(lmc_ctx *c) {
    if (!c) return fail(LMC_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->film, 0, (size_t)c->sc.cam.width * c->sc.cam.height * 3 * sizeof(float), c->stream));
    return LMC_OK;
}
int lmc_film_read(lmc_ctx *c, float *host_rgb) {
    if (!c || !host_rgb) return fail(LMC_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(host_rgb, c->film, (size_t)c->sc.cam.width * c->sc.cam.height * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return LMC_OK;
}
int lmc_film_device_ptr(lmc_ctx *c, void **p) {
    if (!c || !p) return fail(LMC_ERR_ARG, "null argument");
    *p = c->film;
    return LMC_OK;
}
int lmc_film_bind(lmc_ctx *c, void *device_ptr) {
    if (!c || !device_ptr) return fail(LMC_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    if (c->film && c->filmOwned) cudaFree(c->film);
    c->film = (float *)device_ptr; c->filmOwned = false;
    return LMC_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------------------
#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(LMC_ERR_CUDA, std::string(#call) + ": " + nccl_api().GetErrorString(r_)); } while (0)

int lmc_create_multi(const lmc_scene *scene, const int32_t *devices, int32_t n, lmc_ctx **out) {
    if (!scene || !devices || !out || n < 1) return fail(LMC_ERR_ARG, "bad argument");
    for (int i = 0; i < n; i++) out[i] = nullptr;
    for (int i = 0; i < n; i++) {
        const int rc = lmc_create(scene, devices[i], &out[i]);
        if (rc) { for (int k = 0; k < i; k++) { lmc_destroy(out[k]); out[k] = nullptr; } return rc; }
    }
    if (n == 1) return LMC_OK;
    NcclApi &api = nccl_api();
    if (!api.load()) { for (int i = 0; i < n; i++) { lmc_destroy(out[i]); out[i] = nullptr; } return fail(LMC_ERR_UNSUPPORTED, api.error); }
    std::vector<ncclComm_t> comms(n);
    std::vector<int> devs(devices, devices + n);
    const ncclResult_t r = api.CommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) {
        for (int i = 0; i < n; i++) { lmc_destroy(out[i]); out[i] = nullptr; }
        return fail(LMC_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(r));
    }
    for (int i = 0; i < n; i++) out[i]->comm = comms[i];
    return LMC_OK;
}

int lmc_comm_unique_id(void *id128) {
    if (!id128) return fail(LMC_ERR_ARG, "null argument");
    NcclApi &api = nccl_api();
    if (!api.load()) return fail(LMC_ERR_UNSUPPORTED, api.error);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCK(api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return LMC_OK;
}

int lmc_comm_init_rank(lmc_ctx *c, int32_t nranks, int32_t rank, const void *id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(LMC_ERR_ARG, "bad argument");
    if (c->comm) return fail(LMC_ERR_STATE, "this ctx already belongs to a communicator");
    NcclApi &api = nccl_api();
    if (!api.load()) return fail(LMC_ERR_UNSUPPORTED, api.error);
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCK(api.CommInitRank(&c->comm, nranks, id, rank));
    return LMC_OK;
}

int lmc_allreduce_film(lmc_ctx **ctxs, int32_t n) {
    if (!ctxs || n < 1) return fail(LMC_ERR_ARG, "bad argument");
    for (int i = 0; i < n; i++) if (!ctxs[i]) return fail(LMC_ERR_ARG, "null ctx");
    if (n == 1 && !ctxs[0]->comm) return LMC_OK;                      // a lone GPU: the film is already the sum
    for (int i = 0; i < n; i++) if (!ctxs[i]->comm) return fail(LMC_ERR_STATE, "ctx without a communicator (lmc_create_multi / lmc_comm_init_rank)");
    NcclApi &api = nccl_api();
    const size_t count = (size_t)ctxs[0]->sc.cam.width * ctxs[0]->sc.cam.height * 3;
    NCK(api.GroupStart());
    for (int i = 0; i < n; i++) {
        lmc_ctx *c = ctxs[i];
        const ncclResult_t r = api.AllReduce(c->film, c->film, count, ncclFloat32, ncclSum, c->comm, c->stream);
        if (r != ncclSuccess) { api.GroupEnd(); return fail(LMC_ERR_CUDA, std::string("ncclAllReduce: ") + api.GetErrorString(r)); }
    }
    NCK(api.GroupEnd());
    return LMC_OK;
}

// ---- film output (host) ----------------------------------------------------------------------------------------
int lmc_merge_buffer(const float *buffer1, float w1, const float *buffer2, float w2, int64_t n, float *film) {
    if (!film || n < 0) return fail(LMC_ERR_ARG, "bad argument");
    lmc_host::merge_buffer(buffer1, w1, buffer2, w2, (long long)n, film);
    return LMC_OK;
}

int lmc_write_image(const char *path, int32_t width, int32_t height, const float *rgb) {
    if (!path || !rgb || width < 1 || height < 1) return fail(LMC_ERR_ARG, "bad argument");
    if (!lmc_host::write_image(path, width, height, rgb)) return fail(LMC_ERR_IO, std::string("cannot write ") + path);
    return LMC_OK;
}

int32_t lmc_vert_param_size(int32_t cam_depth, int32_t light_depth) {
    const int maxDepth = cam_depth + light_depth;
    return maxDepth * 46 + maxDepth * 10 + 56 + maxDepth * 2 + maxDepth * 1 + 3 + 1 + 1;
}

int lmc_eval_batch(lmc_ctx *c, int32_t cam_depth, int32_t light_depth, int32_t n, const float *lens, const float *primary,
                   const float *vert_params, int32_t vert_stride, float *log_lum, float *grad, float *hess) {
    (void)lens;   // Static mode: the screen position is primary[..], `lens` is unused by the generated code too
    if (!c || !primary || !vert_params || !log_lum || n < 0) return fail(LMC_ERR_ARG, "bad argument");
    if (cam_depth < 1 || light_depth < 0 || cam_depth + light_depth < 3) return fail(LMC_ERR_ARG, "no path function for this (camDepth, lightDepth)");
    const int len = cam_depth + light_depth - 1;
    if (len > 8) return fail(LMC_ERR_UNSUPPORTED, "path functions exist for camDepth + lightDepth - 1 <= 8 (maxDervDepth)");
    if (vert_stride < lmc::serialized_vert_size(cam_depth, light_depth)) return fail(LMC_ERR_ARG, "vert_stride too small for this path class");
    if (n == 0) return LMC_OK;
    CK(cudaSetDevice(c->device));
    const int dim = 2 * (len > 2 ? len : 2);
    if (hess && !grad) return fail(LMC_ERR_ARG, "hess requires grad");
    DevBuf<float> dS, dP, dV, dL, dG, dH;
    if (hess) CK(dH.alloc((size_t)n * dim * dim));
    CK(dS.alloc(38));
    CK(dP.alloc((size_t)n * (dim + 1)));
    CK(dV.alloc((size_t)n * vert_stride));
    CK(dL.alloc((size_t)n));
    if (grad) CK(dG.alloc((size_t)n * dim));
    CK(cudaMemcpyAsync(dS, c->sc.sceneSer, 38 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dP, primary, sizeof(float) * (size_t)n * (dim + 1), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(dV, vert_params, sizeof(float) * (size_t)n * vert_stride, cudaMemcpyHostToDevice, c->stream));
    if (hess) k_eval_batch_hess<<<(n + 63) / 64, 64, 0, c->stream>>>(cam_depth, light_depth, n, dS, dP, dim + 1, dV, vert_stride, dL, dG, dH, dim);
    else k_eval_batch<<<(n + 63) / 64, 64, 0, c->stream>>>(cam_depth, light_depth, n, dS, dP, dim + 1, dV, vert_stride, dL, dG, dim, c->sc.opt.adjointCompat);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(log_lum, dL, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (grad) CK(cudaMemcpyAsync(grad, dG, sizeof(float) * (size_t)n * dim, cudaMemcpyDeviceToHost, c->stream));
    if (hess) CK(cudaMemcpyAsync(hess, dH, sizeof(float) * (size_t)n * dim * dim, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return LMC_OK;
}

int lmc_bvh_probe(lmc_ctx *c, int32_t n, const float *rays, float tmin, float tmax, int32_t any_hit, int32_t *tri_id,
                  int32_t *geom_prim, float *tuv) {
    if (!c || !rays || !tri_id || n < 0) return fail(LMC_ERR_ARG, "bad argument");
    if (n == 0) return LMC_OK;
    CK(cudaSetDevice(c->device));
    DevBuf<float> dR, dT; DevBuf<int> dI, dG;
    CK(dR.alloc(6 * (size_t)n));
    CK(dI.alloc((size_t)n));
    CK(dG.alloc(2 * (size_t)n));
    CK(dT.alloc(3 * (size_t)n));
    CK(cudaMemcpyAsync(dR, rays, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    k_bvh_probe<<<(n + 127) / 128, 128, 0, c->stream>>>(c->sc, n, dR, tmin, tmax, any_hit, dI, dG, dT);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tri_id, dI, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (geom_prim && !any_hit) CK(cudaMemcpyAsync(geom_prim, dG, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (tuv && !any_hit) CK(cudaMemcpyAsync(tuv, dT, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return LMC_OK;
}

}  // extern "C"
