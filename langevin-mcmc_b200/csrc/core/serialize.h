// serialize.h -- gather a path into the reference's flat `primary` / `vertParams` buffers
//  and evaluate its PSS gradient.
//
// Reference: Serialize(scene, path, subPath) src/path.cpp:2497-2586; TriangleMesh::Serialize
// src/trianglemesh.cpp:145-187; BSDF::Serialize lambertian.cpp:10-13, phong.cpp:14-20,
// roughdielectric.cpp:13-20; Light::Serialize pointlight.cpp:15-19, arealight.cpp:17-22,
// envlight.cpp:83-119; the call site src/mutation_mala.h:94-110.
#pragma once
#include "path.h"
#include "pathgrad.h"
#include "pathgrad_rev.h"

namespace lmc {

LMC_HD void st3(float *b, V3 v) { b[0] = v.x; b[1] = v.y; b[2] = v.z; }

// 46 floats
// COMPACT (internal gradient / Hessian path only): skip the fields our evaluator never reads -- the second
// (motion) copy of the triangle, the hasST flag and the texture coordinates, floats 20..44 -- so the
// serialized path costs 25 fewer local-memory stores per vertex.  The layout is unchanged.
template <bool COMPACT = false>
LMC_HD_NOINLINE void serialize_shape(const Scene &sc, int tid, float *b) {
    const TriGeom &tg = sc.tris[tid];
    const TriShade &ts = sc.shade[tid];
    const Material &m = sc.mats[tg.geom];
    b[0] = 0.0f;   // ShapeType::TriangleMesh
    b[1] = 0.0f;   // isMoving
    if (COMPACT) {
        float *q = b + 2;
        for (int k = 0; k < 3; k++) { q[k] = tg.p0[k]; q[3 + k] = tg.e1[k]; q[6 + k] = tg.e2[k]; q[9 + k] = ts.n0[k]; q[12 + k] = ts.n1[k]; q[15 + k] = ts.n2[k]; }
        b[45] = m.invTotalArea;
        return;
    }
    for (int t = 0; t < 2; t++) {
        float *q = b + 2 + 18 * t;
        for (int k = 0; k < 3; k++) { q[k] = tg.p0[k]; q[3 + k] = tg.e1[k]; q[6 + k] = tg.e2[k]; q[9 + k] = ts.n0[k]; q[12 + k] = ts.n1[k]; q[15 + k] = ts.n2[k]; }
    }
    b[38] = m.hasST ? 0.0f : 1.0f;   // sic: 1.0 when the mesh has NO st (src/trianglemesh.cpp:175)
    b[39] = ts.st0[0]; b[40] = ts.st0[1]; b[41] = ts.st1[0]; b[42] = ts.st1[1]; b[43] = ts.st2[0]; b[44] = ts.st2[1];
    b[45] = m.invTotalArea;          // defined only for emitters (the reference leaves totalArea unset otherwise)
}

// 10 floats (padded)
LMC_HD_NOINLINE void serialize_bsdf(const Scene &sc, int tid, V2 st, float *b) {
    const BsdfParams p = bsdf_params(sc, sc.tris[tid].geom, st);
    for (int i = 0; i < LMC_SER_BSDF; i++) b[i] = 0.0f;
    b[0] = (float)p.type;
    if (p.type == BSDF_LAMBERTIAN) { st3(b + 1, p.Kd); }
    else if (p.type == BSDF_PHONG) { st3(b + 1, p.Kd); st3(b + 4, p.Ks); b[7] = p.exponent; b[8] = p.KsWeight; }
    else { st3(b + 1, p.Ks); st3(b + 4, p.Kt); b[7] = p.eta; b[8] = p.invEta; b[9] = p.alpha; }
}

// 56 floats (padded)
template <bool COMPACT = false>
LMC_HD_NOINLINE void serialize_light(const Scene &sc, int light, int lPrimID, float *b) {
    const Light &l = sc.lights[light];
    // (COMPACT: an env-light record is written completely below -- entries 0..53 are all the evaluator reads)
    if (!(COMPACT && l.type == LIGHT_ENV)) for (int i = 0; i < LMC_SER_LIGHT; i++) b[i] = 0.0f;
    b[0] = (float)l.type;
    if (l.type == LIGHT_POINT) { for (int k = 0; k < 3; k++) { b[1 + k] = l.pos[k]; b[4 + k] = l.emission[k]; } }
    else if (l.type == LIGHT_AREA) {
        serialize_shape<COMPACT>(sc, light_prim_tid(sc, l, lPrimID), b + 1);
        for (int k = 0; k < 3; k++) b[1 + LMC_SER_SHAPE + k] = l.emission[k];
    } else {
        const EnvMap &e = sc.env;
        float *q = b + 1;
        for (int k = 0; k < 15; k++) { q[k] = e.toWorldSer[k]; q[15 + k] = e.toLightSer[k]; }
        q += 30;
        const int col = lPrimID % e.width, row = lPrimID / e.width;
        const float *cdfCol = e.cdfCols + row * (e.width + 1);
        q[0] = cdfCol[col]; q[1] = cdfCol[col + 1]; q[2] = e.cdfRows[row]; q[3] = e.cdfRows[row + 1];
        q[4] = (float)col; q[5] = (float)row; q[6] = e.pixelSize[0]; q[7] = e.pixelSize[1];
        st3(q + 8, env_rep_at(e, col, row)); st3(q + 11, env_rep_at(e, col + 1, row));
        st3(q + 14, env_rep_at(e, col, row + 1)); st3(q + 17, env_rep_at(e, col + 1, row + 1));
        q[20] = e.rowWeights[dm_clampi(row, 0, e.height - 1)];
        q[21] = e.rowWeights[dm_clampi(row + 1, 0, e.height - 1)];
        q[22] = e.normalization;
    }
}

// upper bound on the vertParams floats serialize_path writes for a path of (c, l)
LMC_HD int serialized_vert_size(int camDepth, int lgtDepth) {
    const int nl = lgtDepth > 1 ? lgtDepth - 1 : 0, nc = camDepth > 1 ? camDepth - 1 : 0;
    return 3 + 1 + LMC_SER_LIGHT + (nl + nc) * (LMC_SER_SHAPE + 2 + LMC_SER_BSDF + 1) + LMC_SER_LIGHT + LMC_SER_BSDF + 1;
}

// Serialize(scene, path, subPath): returns the number of vertParams floats written
template <int MAXD, bool COMPACT = false>
LMC_HD_NOINLINE int serialize_path(const Scene &sc, const Path<MAXD> &path, float *primary, float *vertParams) {
    int pi = 0;
    primary[pi++] = path.time;
    float *b = vertParams;
    st3(b, path.lensVertexPos); b += 3;
    if (path.lgtDepth > 1) {
        primary[pi++] = path.lgtRndPos.x; primary[pi++] = path.lgtRndPos.y;
        primary[pi++] = path.lgtRndDir.x; primary[pi++] = path.lgtRndDir.y;
        *b++ = pick_light_prob(sc, path.lgtLight);
        serialize_light<COMPACT>(sc, path.lgtLight, path.lgtPrim, b); b += LMC_SER_LIGHT;
        for (int d = 0; d < path.nLgt; d++) {
            const SurfaceVertex &sv = path.lgt[d];
            serialize_shape<COMPACT>(sc, sv.tid, b); b += LMC_SER_SHAPE;
            *b++ = sv.bsdfDiscrete; *b++ = sv.useAbsoluteParam;
            serialize_bsdf(sc, sv.tid, sv.st, b); b += LMC_SER_BSDF;
            if (d == path.nLgt - 1 && path.camDepth == 1) return (int)(b - vertParams);
            if (d == path.nLgt - 1) break;
            primary[pi++] = sv.bsdfRndParam.x; primary[pi++] = sv.bsdfRndParam.y;
            *b++ = sv.rrWeight;
        }
    }
    primary[pi++] = path.screenPos.x; primary[pi++] = path.screenPos.y;
    for (int d = 0; d < path.nCam; d++) {
        const SurfaceVertex &sv = path.cam[d];
        if (sv.tid >= 0) serialize_shape<COMPACT>(sc, sv.tid, b);
        else { for (int i = 0; i < (COMPACT ? 20 : LMC_SER_SHAPE); i++) b[i] = 0.0f; b[45] = 0.0f; }
        b += LMC_SER_SHAPE;
        if (d == path.nCam - 1) {
            if (path.lgtDepth == 0) {
                if (path.envLight >= 0) {
                    serialize_light<COMPACT>(sc, path.envLight, path.envPrim, b); b += LMC_SER_LIGHT;
                    *b++ = pick_light_prob(sc, path.envLight);
                } else {
                    const TriGeom &tg = sc.tris[sv.tid];
                    const int light = sc.mats[tg.geom].areaLight;
                    serialize_light<COMPACT>(sc, light, tg.prim, b); b += LMC_SER_LIGHT;
                    *b++ = pick_light_prob(sc, light);
                }
            } else if (path.lgtDepth == 1) {
                primary[pi++] = sv.dlRndParam.x; primary[pi++] = sv.dlRndParam.y;
                serialize_light<COMPACT>(sc, sv.dlLight, sv.dlPrim, b); b += LMC_SER_LIGHT;
                serialize_bsdf(sc, sv.tid, sv.st, b); b += LMC_SER_BSDF;
                *b++ = pick_light_prob(sc, sv.dlLight);
            } else {
                serialize_bsdf(sc, sv.tid, sv.st, b); b += LMC_SER_BSDF;
            }
            return (int)(b - vertParams);
        }
        primary[pi++] = sv.bsdfRndParam.x; primary[pi++] = sv.bsdfRndParam.y;
        *b++ = sv.bsdfDiscrete; *b++ = sv.useAbsoluteParam;
        serialize_bsdf(sc, sv.tid, sv.st, b); b += LMC_SER_BSDF;
        *b++ = sv.rrWeight;
    }
    return (int)(b - vertParams);
}

// The reference has a derivative function for every (c, l) with c >= 1, c + l >= 3 and
// c + l - 1 <= maxDervDepth (src/path.cpp:3955-3959); MLT paths always have c + l >= 3.
template <int MAXD>
LMC_HD bool grad_supported(const Scene &sc, const Path<MAXD> &path) {
    const int len = path.camDepth + path.lgtDepth - 1;
    return path.camDepth >= 1 && path.camDepth + path.lgtDepth > 2 && len <= sc.opt.maxDervDepth;
}

// Serialized-path scratch large enough for every path the PSS_MAX_LENGTH gate lets through
// (LMC: dim <= pssMaxLength <= 12; H2MC: c + l - 1 <= maxDervDepth = 8  =>  at most 8 surface vertices).
#define LMC_GRAD_MAX_SURF 8
#define LMC_GRAD_SCRATCH (3 + 1 + LMC_SER_LIGHT + LMC_GRAD_MAX_SURF * (LMC_SER_SHAPE + 2 + LMC_SER_BSDF + 1) + LMC_SER_LIGHT + LMC_SER_BSDF + 1)

#if defined(LMC_ORACLE_HOOK) && !defined(__CUDA_ARCH__)
// Test-only hook (oracle build): lets "Oracle-R" route the gradient through the REFERENCE's own
// generated reverse-mode code (oracle/_ref) instead of the evaluator below.  Never compiled into
// the product.
typedef void (*RefGradHook)(int camDepth, int lightDepth, const float *lens, const float *primary, const float *sceneSer,
                            float *vertParams, int nVert, float *grad, int dim);
inline RefGradHook &ref_grad_hook() { static thread_local RefGradHook h = nullptr; return h; }
#endif

// dervFunc(screenPos, primary, sceneParams, vertParams, vGrad, NULL) for the path's (c, l)
template <int MAXD>
LMC_HD_NOINLINE void path_gradient(const Scene &sc, const Path<MAXD> &path, float *grad) {
    float primary[2 * LMC_GRAD_MAX_SURF + 1 + 4];
    float vertParams[LMC_GRAD_SCRATCH];
#if defined(LMC_ORACLE_HOOK) && !defined(__CUDA_ARCH__)
    if (ref_grad_hook()) {
        const int dim_ = path_dimension(path);
        float big[1100];
        for (int i = 0; i < 1100; i++) big[i] = 0.0f;
        float prim_[32];
        const int nv = serialize_path(sc, path, prim_, big);
        if (path.lgtDepth == 0 && path.envLight >= 0 && path.nCam > 0) {
            // The reference leaves the shape block of an environment hit untouched (src/path.cpp:2545-2548): its code
            // then differentiates whatever the reused buffer held.  Emulate "stale but finite" with a fixed triangle
            // (a zero block would turn the whole reverse sweep into NaN, which the reference only sees on a chain's
            // very first evaluation).
            float *q = big + 3 + (path.nCam - 1) * (LMC_SER_SHAPE + 2 + LMC_SER_BSDF + 1) + 2;
            const float tri[18] = {-1e3f, -1e3f, 977.0f, 2e3f, 0.0f, 13.0f, 0.0f, 2e3f, -7.0f, 0, 0, 1, 0, 0, 1, 0, 0, 1};
            for (int i = 0; i < 18; i++) q[i] = tri[i];
        }
        const float lens[2] = {path.screenPos.x, path.screenPos.y};
        ref_grad_hook()(path.camDepth, path.lgtDepth, lens, prim_, sc.sceneSer, big, nv, grad, dim_);
        return;
    }
#endif
    const int dim = path_dimension(path);
    const int nSurf = (path.camDepth > 1 ? path.camDepth - 1 : 0) + (path.lgtDepth > 1 ? path.lgtDepth - 1 : 0);
    if (nSurf > LMC_GRAD_MAX_SURF || dim > 2 * LMC_GRAD_MAX_SURF) {   // outside the gate: caller should not ask
        for (int i = 0; i < dim; i++) grad[i] = 0.0f;
        return;
    }
    serialize_path<MAXD, true>(sc, path, primary, vertParams);
    path_loglum_grad_mode(sc.opt.adjointCompat, path.camDepth, path.lgtDepth, sc.sceneSer, primary, vertParams, grad);
}

// dervFunc(..., vGrad, vHess) of the H2MC library (src/mutation_h2mc.h:74-79); hess row-major dim x dim
template <int MAXD>
LMC_HD_NOINLINE void path_hessian(const Scene &sc, const Path<MAXD> &path, float *grad, float *hess) {
    float primary[2 * LMC_GRAD_MAX_SURF + 1 + 4];
    float vertParams[LMC_GRAD_SCRATCH];
    const int dim = path_dimension(path);
    const int nSurf = (path.camDepth > 1 ? path.camDepth - 1 : 0) + (path.lgtDepth > 1 ? path.lgtDepth - 1 : 0);
    if (nSurf > LMC_GRAD_MAX_SURF || dim > LMC_HESS_MAXDIM) {
        for (int i = 0; i < dim; i++) grad[i] = 0.0f;
        for (int i = 0; i < dim * dim && i < LMC_HESS_MAXDIM * LMC_HESS_MAXDIM; i++) hess[i] = 0.0f;
        return;
    }
    serialize_path<MAXD, true>(sc, path, primary, vertParams);
#if LMC_HESS_REV_CHUNK > 0
    path_loglum_hess_rev<LMC_HESS_REV_CHUNK>(path.camDepth, path.lgtDepth, sc.sceneSer, primary, vertParams, grad, hess);
#else
    path_loglum_hess(path.camDepth, path.lgtDepth, sc.sceneSer, primary, vertParams, grad, hess);
#endif
}

}  // namespace lmc
