// scene.h -- flattened, pointer-based scene view shared by host twin and device kernels.
//
// This is the HBM layout of everything the chain kernel reads (DESIGN.md "Data layout"):
// triangles are stored in BVH order ("tid" = position in that order) so a leaf is a
// contiguous run and the id a ray query returns indexes every per-triangle array directly.
// It replaces the reference's pointer graph Scene -> Shape -> TriMeshData / BSDF / Light
// (src/scene.h:23-45, src/trianglemesh.h:22-31, src/bsdf.h:10-68, src/light.h:14-59).
#pragma once
#include "vec.h"

namespace lmc {

enum BsdfType { BSDF_LAMBERTIAN = 0, BSDF_PHONG = 1, BSDF_ROUGHDIELECTRIC = 2 };  // src/bsdf.h:6
enum LightType { LIGHT_POINT = 0, LIGHT_AREA = 1, LIGHT_ENV = 2 };                // src/light.h:7

// 48 B, read as 3 x float4 during traversal
struct TriGeom {
    float p0[3]; int geom;      // geom = shape index (the reference's geomID)
    float e1[3]; int prim;      // prim = index inside that shape (the reference's primID)
    float e2[3]; int pad;
};
// 64 B, read once per accepted hit
struct TriShade {
    float n0[3], n1[3], n2[3];
    float st0[2], st1[2], st2[2];
    int pad;
};
// 64 B BVH2 node: both child boxes inline.  child >= 0: inner node index;
// child < 0: leaf, ~child = (first << 3) | (count - 1).
struct BvhNode {
    float lmin[3], lmax[3];
    float rmin[3], rmax[3];
    int left, right;
    int pad[2];
};

struct Material {
    int type;          // BsdfType
    int twoSided;
    int kdTex;         // texture index for Kd (Lambertian/Phong diffuse) or -1
    int areaLight;     // light index if the shape is an emitter, else -1
    float Kd[3];
    float Ks[3];
    float Kt[3];
    float exponent;
    float KsWeight;    // src/phong.cpp:159-169
    float eta, invEta;
    float alpha;
    int hasST;         // mesh has texture coordinates (else st = barycentric uv)
    float invTotalArea;
    int firstTid;      // not used by traversal; kept for diagnostics
    // bitmap-textured parameters (src/parsescene.cpp:341-412 Parse3DMap / Parse1DMap), -1 = the constant above:
    // Phong specularReflectance / exponent, RoughDielectric specularReflectance / specularTransmittance / alpha.
    // One-channel maps read the texture's first channel (BitmapTexture<1>, src/bitmaptexture.h:73-97).
    int ksTex, ktTex, expTex, alphaTex;
};

struct Texture {
    int width, height;
    int offset;        // float offset into texData (RGB interleaved)
    float gamma;       // 2.2 for 8-bit sources, 1 otherwise (src/bitmaptexture.h:136-144)
    float sScale, tScale;
};

struct Light {
    int type;          // LightType
    float samplingWeight;
    // area light
    int geom;          // shape index
    int numPrims;
    int primCdfOffset; // offset into lightCdf: numPrims+1 floats (PiecewiseConstant1D cdf)
    int primTidOffset; // offset into lightPrimTid: numPrims ints (prim -> tid)
    float emission[3];
    float invTotalArea;
    // point light
    float pos[3];
};

struct EnvMap {
    int present;
    int lightIndex;
    int width, height;
    const float *image;      // width*height*3
    const float *cdfRows;    // height+1
    const float *cdfCols;    // (width+1)*height
    const float *rowWeights; // height
    float normalization;
    float pixelSize[2];
    M44 toWorld, toLight;    // static transforms (Interpolate(...) of a non-moving transform)
    float toWorldSer[15], toLightSer[15];  // AnimatedTransform serialisation (for the reference ABI)
};

struct Camera {
    M44 sampleToCam, camToSample;
    M44 camToWorld, worldToCam;
    float nearClip, farClip;
    float dist;
    int width, height;
    float camToWorldSer[15];
};

struct Options {                 // src/dptoptions.h:7-34 + compile-time constants
    int minDepth, maxDepth;
    int bidirectional;
    int h2mc, mala;
    int numChains;
    int seedOffset;
    int useLightCoordinateSampling;
    int largeStepMultiplexed;
    int cacheEnabled;            // option `globalcache`: the cross-chain cache of src/global_cache.h (default 0: every
                                 // eligible MALA step evaluates its gradient -- the benchmark / parity mode)
    int maxDervDepth;            // 8
    int pssMinLength, pssMaxLength;   // 2, 12
    int adjointCompat;           // gradient of the mutations: 1 reverse sweep in the reference's merge order (default),
                                 // 0 reverse sweep / true adjoint, 2 forward-mode duals (core/pathgrad_rev.h)
    int outlierWeakRejectCnt;    // OUTLIER_WEAK_REJECT_CNT 10000    (src/mutation.h:6)
    int outlierStrongRejectCnt;  // OUTLIER_STRONG_REJECT_CNT 1000   (src/mutation.h:7)
    float outlierRatioThreshold; // OUTLIER_RATIO_THRESHOLD 30       (src/mutation.h:8)
    float perturbStdDev;         // 0.01
    float roughnessThreshold;    // 0.05
    float largeStepProbability;  // 0.05
    float largeStepProbScale;    // 1 (4 in the shipped LMC xml)
    float malaGN;                // 100
    float malaStepsize;          // 0.005
    float malaStdDev;            // 0.005
    float discreteStdDev;        // 0.01
    float uniformMixingProbability;  // 0.1
    float lsRatio;               // LS_RATIO 0.1
    // H2MCParam (src/h2mc.h:9-23), evaluated once on the host from sigma = perturbStdDev, L = pi/2
    float h2mcL, h2mcPosScale, h2mcPosOffset, h2mcNegScale, h2mcNegOffset;
};

// Global cache of adaptation states (src/global_cache.h:16-163, instantiated at src/mlt.cpp:53): for every PSS
// dimension D in [PSS_MIN_LENGTH, PSS_MAX_LENGTH] up to PSS_MAX_SIZE = 3000 entries (pss[D], v1[D], v2[D]) pushed by
// chains that leave a MALA-adapted state with an accepted large step (src/mlt.cpp:121-127).  Once a dimension is
// full, chains of that dimension stop evaluating gradients and look their moments up here
// (src/mutation_mala.h:131-161).  D = 2 * max(c + l - 1, 2) is even and >= 4, so five slots cover D = 4 .. 12.
#define LMC_CACHE_SLOTS 5
#define LMC_CACHE_MAX_SIZE 3000          // PSS_MAX_SIZE
#define LMC_CACHE_QUERY_DIST 0.01f       // PSS_QUERY_DIST
#define LMC_CACHE_REUSE_DIST 0.10f       // PSS_REUSE_DIST
#define LMC_CACHE_KNN 5
// A ready slot is read-only (global_cache_t::push refuses once is_ready), so its 3000 entries are binned ONCE into a
// uniform grid over the first three PSS coordinates: cell size 1/24 >= the largest query radius 0.01 * sqrt(12), hence
// every entry within the radius of a query lies in the 3 x 3 x 3 cells around the query's cell and a query tests ~6
// entries instead of 3000 (the reference walks a KD-tree, nanoflann).  Same result as the linear scan by construction.
#define LMC_CACHE_GRID 24
#define LMC_CACHE_CELLS (LMC_CACHE_GRID * LMC_CACHE_GRID * LMC_CACHE_GRID)
#define LMC_CACHE_GRID_INTS (2 * LMC_CACHE_CELLS + 1 + LMC_CACHE_MAX_SIZE)     // cellStart[CELLS + 1], cursor[CELLS], entry[MAX_SIZE]
struct GlobalCacheView {
    float *data;     // slot s (D = 4 + 2 s) starts at cache_slot_offset(s); entry e = 3 D floats: pss, v1, v2
    int *count;      // [LMC_CACHE_SLOTS] entries stored
    int *ready;      // [LMC_CACHE_SLOTS] is_ready
    int *grid;       // [LMC_CACHE_SLOTS][LMC_CACHE_GRID_INTS] (may be null: linear scan)
    int *gridReady;  // [LMC_CACHE_SLOTS] the slot's grid has been built
};
LMC_HD int cache_cell_coord(float x) {
    const int c = (int)(x * (float)LMC_CACHE_GRID);
    return c < 0 ? 0 : (c > LMC_CACHE_GRID - 1 ? LMC_CACHE_GRID - 1 : c);
}
LMC_HD int cache_cell(const float *pss) {
    return (cache_cell_coord(pss[0]) * LMC_CACHE_GRID + cache_cell_coord(pss[1])) * LMC_CACHE_GRID + cache_cell_coord(pss[2]);
}
LMC_HD int cache_slot(int dim) { return (dim >= 4 && dim <= 12 && (dim & 1) == 0) ? (dim - 4) / 2 : -1; }
LMC_HD int cache_slot_offset(int s) {        // floats before slot s: 3000 * 3 * sum_{k<s} (4 + 2k)
    return LMC_CACHE_MAX_SIZE * 3 * (4 * s + s * (s - 1));
}
#define LMC_CACHE_FLOATS (LMC_CACHE_MAX_SIZE * 3 * (4 + 6 + 8 + 10 + 12))

struct Scene {
    // geometry
    int numTris, numNodes, numGeoms;
    const TriGeom *tris;
    const TriShade *shade;
    const BvhNode *nodes;
    const Material *mats;        // per geom
    // textures
    int numTextures;
    const Texture *textures;
    const float *texData;
    // lights
    int numLights;
    const Light *lights;
    const float *lightPickCdf;   // numLights+1 (PiecewiseConstant1D over samplingWeight)
    float lightWeightSum;
    const float *lightCdf;
    const int *lightPrimTid;
    EnvMap env;
    Camera cam;
    float bsphereCenter[3];
    float bsphereRadius;         // already x1000 (src/scene.cpp:40)
    float sceneSer[38];          // Serialize(scene) for the reference ABI (src/scene.cpp:164-169)
    Options opt;
    GlobalCacheView gc;          // valid only when opt.cacheEnabled (owned by the ctx / the oracle run)
};

#define LMC_ISECT_EPS 5e-4f   // c_IsectEpsilon, src/commondef.h:53
#define LMC_SHADOW_EPS 5e-4f  // c_ShadowEpsilon, src/commondef.h:54
#define LMC_COS_EPS 1e-4f     // c_CosEpsilon, src/commondef.h:60

LMC_HD V3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

}  // namespace lmc
