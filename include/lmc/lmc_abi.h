/* lmc_abi.h -- thin C ABI of the B200-native Langevin-MCMC chain sampler (liblmc_b200.so).
 *
 * Plain pointers and sizes only; no C++ / torch types cross this boundary.  Every entry point
 * returns 0 on success and a negative code on failure (never throws); the message of the last
 * failure on the calling thread is available from lmc_last_error().  There is NO CPU fallback:
 * every compute entry point fails with LMC_ERR_CUDA when no sm_100 device is usable.
 *
 * What each entry point replaces in the reference (luanfujun/Langevin-MCMC):
 *
 *   lmc_scene_load / _free / options     ParseScene(filename)            src/parsescene.h:8, src/parsescene.cpp:627-639
 *                                        DptOptions                      src/dptoptions.h:7-34
 *   lmc_mlt_init / lmc_mlt_init_device   MLTInit(...)                    src/mlt.h:41-154 (phase before the loop; host / device paths)
 *   lmc_create / lmc_destroy             Scene::Scene (Embree build)     src/scene.cpp:8-46, src/trianglemesh.cpp:107-143
 *   lmc_chains_begin + lmc_run_chains    the ParallelFor chain lambda    src/mlt.cpp:60-196, called with
 *                                        Mutation::Mutate plugins        src/mutation.h:16-26
 *   lmc_film_* / lmc_stats               SampleBuffer indirectBuffer     src/mlt.cpp:55, src/image.h:54-77
 *   lmc_direct_lighting                  DirectLighting(scene, buffer)   src/direct.cpp:4-54 (pre-pass of MLT(), src/mlt.cpp:30-34)
 *   lmc_eval_batch                       PathFunc / PathFuncDerv         src/path.h:121-125 (dlsym'd
 *                                        evaluate_path_bidir_mala_<c>_<l>_static[_derv], src/path.cpp:3389-3417)
 *   lmc_bvh_probe                        Intersect / Occluded            src/scene.cpp:106-149 (rtcIntersect1 / rtcOccluded1)
 *   lmc_merge_buffer / lmc_write_image   MergeBuffer, BufferToFilm,      src/image.h:79-105, src/image.cpp:29-60
 *                                        WriteImage                      (end of MLT(), src/mlt.cpp:203-212, and the
 *                                                                        progressive dump, src/mlt.cpp:171-193)
 *   lmc_create_multi / lmc_comm_* /      the shared SampleBuffer all     src/mlt.cpp:55 (indirectBuffer), src/image.h:66-77;
 *   lmc_allreduce_film                   chain threads splat into        one ctx per GPU + one NCCL all-reduce
 *
 * C++ mirror of the reference's plugin-surface headers (same type and field names) over this ABI:
 * include/lmc/{dptoptions,parsescene,bsdf,mutation,mlt}.h.
 */
#ifndef LMC_ABI_H
#define LMC_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMC_OK 0
#define LMC_ERR_ARG (-1)
#define LMC_ERR_IO (-2)
#define LMC_ERR_CUDA (-3)
#define LMC_ERR_STATE (-4)
#define LMC_ERR_UNSUPPORTED (-5)

typedef struct lmc_scene lmc_scene; /* host-side parsed + flattened scene (parsescene.h surface) */
typedef struct lmc_ctx lmc_ctx;     /* one GPU: scene in HBM + chain state + film */

/* integer facts about a loaded scene */
typedef struct lmc_scene_info {
    int32_t width, height;       /* film */
    int32_t num_triangles, num_bvh_nodes, num_lights, num_shapes, num_textures;
    int32_t spp, direct_spp, num_init_samples; /* <dpt> spp / directspp / numinitsamples */
    int32_t report_interval_spp;               /* <dpt> reportintervalspp (0 = no progressive dumps) */
} lmc_scene_info;

/* counters accumulated by lmc_run_chains (sum over the ctx's chains since lmc_chains_begin) */
typedef struct lmc_stats {
    uint64_t proposed[4]; /* indexed by MutationType {Large, Small, H2MCSmall, MALASmall}, src/mutation.h:11 */
    uint64_t accepted[4];
    uint64_t gradient_evals;     /* PSS-gradient evaluations (the reference's dervFunc calls) */
    uint64_t gradient_nonfinite; /* gradients zeroed by the IsFinite guard, src/mutation_mala.h:108-110 */
    uint64_t kernel_launches;    /* CUDA kernels launched by this ctx so far */
    double last_kernel_ms;       /* device time of the chain kernels of the last lmc_run_chains (CUDA events) */
    uint64_t outlier_resets;     /* chain resets of the "stuck chain" rule, src/mlt.cpp:147-169 */
    uint64_t cache_queries;      /* global cache (option globalcache): global_cache_t::query calls, src/global_cache.h:96 */
    uint64_t cache_hits;         /* ... that found a neighbour within PSS_QUERY_DIST */
    uint32_t cache_count[5];     /* entries stored per PSS dimension D = 4, 6, 8, 10, 12 (3000 = ready) */
    uint32_t reserved_u32;
} lmc_stats;

/* chain-run descriptor: the loop-invariant inputs of the lambda at src/mlt.cpp:60-90 */
typedef struct lmc_run_desc {
    int32_t num_chains;          /* chains owned by THIS ctx */
    int32_t chain_base;          /* global id of this ctx's first chain (RNG seed = id + seedoffset) */
    int32_t total_chains;        /* numChains of the whole job (outlier-reset walk, src/mlt.cpp:158) */
    int32_t reserved;
    int64_t samples_per_chain;   /* numSamplesThisChain (only the LS_RATIO schedule reads it, src/mlt.cpp:96) */
    float normalization;         /* avgScore from MLTInit */
    float reserved_f;
} lmc_run_desc;

const char *lmc_last_error(void);
const char *lmc_version(void);

/* ---- scene (host) ------------------------------------------------------------------------- */
/* path: reference-format scene .xml (images pre-decoded by tools/stage_scenes.py) or a .pack */
int lmc_scene_load(const char *path, lmc_scene **out);
int lmc_scene_save_pack(const lmc_scene *scene, const char *path);
void lmc_scene_free(lmc_scene *scene);
int lmc_scene_get_info(const lmc_scene *scene, lmc_scene_info *out);
/* options by their <dpt> name ("maxdepth", "mala", "largestepprob", ...; see host_scene.h) */
int lmc_scene_set_option(lmc_scene *scene, const char *name, double value);
int lmc_scene_get_option(const lmc_scene *scene, const char *name, double *value);
/* Serialize(scene) -- the 38 floats the reference hands to its path functions (src/scene.cpp:164-169) */
int lmc_scene_serialized(const lmc_scene *scene, float *out38);

/* MLTInit on the host: normalization (= avgScore) and per-chain init lsScore (init_ls_score may
 * be NULL).  logical_threads fixes the RNG streams independent of the machine (reference:
 * NumSystemCores()); 32 reproduces the reference machine. */
int lmc_mlt_init(const lmc_scene *scene, int64_t num_init_samples, int32_t num_chains, int32_t logical_threads,
                 float *normalization, float *init_ls_score);

/* ---- device ------------------------------------------------------------------------------- */
/* Uploads the flattened scene + BVH2 to HBM on CUDA device `device`.  The scene's options are
 * captured at this point. */
int lmc_create(const lmc_scene *scene, int32_t device, lmc_ctx **out);
void lmc_destroy(lmc_ctx *ctx);
/* cudaStream_t to launch on (default: the legacy default stream); pass torch's current stream */
int lmc_set_stream(lmc_ctx *ctx, void *cuda_stream);

/* MLTInit with the init paths generated on the device: logical thread t = one CUDA
 * thread running GeneratePathBidir for its share of the samples (src/mlt.h:51-99); the score sum, CDF and
 * equal-spaced seeding (src/mlt.h:107-153) stay sequential on the host.  Same results, bit for bit, as
 * lmc_mlt_init with the same logical_threads; use a large thread count (e.g. 65536) to fill the GPU. */
int lmc_mlt_init_device(lmc_ctx *ctx, int64_t num_init_samples, int32_t num_chains, int32_t logical_threads,
                        float *normalization, float *init_ls_score);

/* The same, split for several GPUs: every rank generates the init paths of ITS logical threads
 * [thread_begin, thread_end) and gets their lsScores in (thread, sample, contribution) order; the host concatenates
 * the parts in thread order (an all-gather) and runs the sequential tail (score sum, CDF, equal-spaced seeding,
 * src/mlt.h:107-153) with lmc_mlt_init_finish.  Bit-identical to lmc_mlt_init_device / lmc_mlt_init for the same
 * logical_threads.  Call with scores == NULL first to learn *num_scores (the part is generated then and kept), then
 * with a buffer of that capacity. */
int lmc_mlt_init_device_part(lmc_ctx *ctx, int64_t num_init_samples, int32_t logical_threads, int32_t thread_begin, int32_t thread_end,
                             float *scores, int64_t capacity, int64_t *num_scores);
int lmc_mlt_init_finish(const float *scores, int64_t num_scores, int64_t num_init_samples, int32_t num_chains,
                        float *normalization, float *init_ls_score);

/* DirectLighting(scene, buffer) (src/direct.cpp:4-54): the direct-illumination pre-pass of
 * MLT(), direct_spp samples per pixel (the scene's <integer name="directspp">), one RNG per 16 x 16 tile as in
 * the reference.  host_rgb receives the UNWEIGHTED sample buffer (W*H*3 floats); the caller merges it with the
 * chain film as MergeBuffer does: direct / directSpp + indirect / spp (src/mlt.cpp:203-207). */
int lmc_direct_lighting(lmc_ctx *ctx, int32_t direct_spp, float *host_rgb);

/* Allocate + initialise chain state for desc->num_chains chains and clear the film.
 * init_ls_score: HOST pointer to total_chains floats (or NULL = zeros). */
int lmc_chains_begin(lmc_ctx *ctx, const lmc_run_desc *desc, const float *init_ls_score);
/* Advance every chain by `num_mutations` iterations of src/mlt.cpp:91-170 (asynchronous on the
 * ctx stream unless a trace is requested).  trace / a_trace: optional HOST buffers of
 * num_chains x num_mutations entries, [chain][step]: trace byte = mutationType | accepted<<2 |
 * (a>0)<<3, a_trace = acceptance probability. */
int lmc_run_chains(lmc_ctx *ctx, int64_t num_mutations, uint8_t *trace, float *a_trace);
int lmc_synchronize(lmc_ctx *ctx);
int lmc_get_stats(lmc_ctx *ctx, lmc_stats *out);

/* Film = indirectBuffer: W*H*3 fp32 sums of splats (divide by spp as src/mlt.cpp:203-207 does). */
int lmc_film_clear(lmc_ctx *ctx);
int lmc_film_read(lmc_ctx *ctx, float *host_rgb);          /* D2H copy, synchronises */
int lmc_film_device_ptr(lmc_ctx *ctx, void **device_ptr);  /* for an NCCL all-reduce by the caller */
/* Use caller-owned device memory (W*H*3 floats) as the film, e.g. a torch tensor */
int lmc_film_bind(lmc_ctx *ctx, void *device_ptr);

/* ---- multi-GPU: one ctx per device, chains sharded by global id, ONE collective --------------------------------- */
/* One process, n devices: creates a ctx on each and an NCCL communicator over them (ncclCommInitAll).  out[n]. */
int lmc_create_multi(const lmc_scene *scene, const int32_t *devices, int32_t n, lmc_ctx **out);
/* One process per GPU: rank 0 calls lmc_comm_unique_id (128 bytes, an ncclUniqueId), hands it to the other ranks
 * by any side channel (MPI, torch.distributed, a file); every rank then calls lmc_comm_init_rank on its ctx. */
int lmc_comm_unique_id(void *id128);
int lmc_comm_init_rank(lmc_ctx *ctx, int32_t nranks, int32_t rank, const void *id128);
/* Sum of the films of a job, in place, asynchronous on each ctx's stream (ncclAllReduce, fp32, sum).  ctxs[n]: all
 * ctx of THIS process that belong to the communicator (n = 1 with lmc_comm_init_rank, n = the device count with
 * lmc_create_multi).  Statistics are not reduced: add lmc_get_stats over the ctx / ranks. */
int lmc_allreduce_film(lmc_ctx **ctxs, int32_t n);

/* ---- film output (host) ----------------------------------------------------------------------------------------- */
/* film[i] = w1 * buffer1[i] + w2 * buffer2[i] for n floats: MergeBuffer + BufferToFilm (src/image.h:79-105), called
 * by MLT() with (direct, 1/directSpp, indirect, 1/spp) (src/mlt.cpp:203-207).  Either buffer may be NULL (= zeros). */
int lmc_merge_buffer(const float *buffer1, float w1, const float *buffer2, float w2, int64_t n, float *film);
/* WriteImage (src/image.cpp:29-60): rgb = H x W x 3 floats, row 0 on top.  ".exr" -> OpenEXR scan-line file with
 * three uncompressed 32-bit float channels; ".pfm" -> portable float map. */
int lmc_write_image(const char *path, int32_t width, int32_t height, const float *rgb);

/* ---- fine-grained boundary (parity harness) ------------------------------------------------- */
/* n serialized paths of class (cam_depth, light_depth) in the reference's buffer layout
 * (SURVEY.md App. A.4): lens n x 2, primary n x (D+1) [time first], vert_params n x
 * vert_stride, scene = lmc_scene_serialized().  HOST pointers.  Outputs: log_lum[n] =
 * log(Luminance(contrib)); grad n x D (or NULL); hess n x D x D row-major (or NULL; needs grad, D <= 16) --
 * the vGrad / vHess of PathFuncDerv (src/path.h:122-123, src/mutation_h2mc.h:74-79). */
int lmc_eval_batch(lmc_ctx *ctx, int32_t cam_depth, int32_t light_depth, int32_t n, const float *lens,
                   const float *primary, const float *vert_params, int32_t vert_stride, float *log_lum, float *grad,
                   float *hess);
/* size of one vert_params record for class (c, l): GetVertParamSize, src/path.cpp:2485-2495 */
int32_t lmc_vert_param_size(int32_t cam_depth, int32_t light_depth);

/* n rays (org xyz, dir xyz), HOST pointers.  any_hit = 0: closest hit -> tri_id[n] (BVH-order
 * id, -1 = miss), geom_prim[n x 2] = (geomID, primID), tuv[n x 3].  any_hit = 1: tri_id[n] = 1/0
 * occluded flag on [tmin, tmax]. */
int lmc_bvh_probe(lmc_ctx *ctx, int32_t n, const float *rays, float tmin, float tmax, int32_t any_hit,
                  int32_t *tri_id, int32_t *geom_prim, float *tuv);

#ifdef __cplusplus
}
#endif
#endif /* LMC_ABI_H */
