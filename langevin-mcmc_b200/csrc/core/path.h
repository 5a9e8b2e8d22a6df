// path.h -- bidirectional path construction (large step), PSS perturbation replay (small step)
// and path bookkeeping, re-stated from the reference's src/path.cpp.
//
//   BidirPathState                src/path.cpp:529-540
//   EmitFromCamera[Init]          src/path.cpp:542-586
//   EmitFromLight[Init]           src/path.cpp:588-618 (init :574-586)
//   ConvertMIS                    src/path.cpp:620-631
//   ConnectToCamera               src/path.cpp:633-745
//   BSDFSampling<adjoint,perturb> src/path.cpp:747-900
//   HandleHitLight                src/path.cpp:902-967
//   DirectLighting[Init]          src/path.cpp:969-1089 (init :184-193)
//   ConnectVertex                 src/path.cpp:1091-1235
//   GeneratePathBidir             src/path.cpp:1237-1449
//   RussianRoulette               src/path.cpp:388-404
//   ToSubpath                     src/path.cpp:1660-1669
//   PerturbPathBidir              src/path.cpp:1953-2160
//   GetPrimaryParamSize/GetDimension  src/path.cpp:2481-2483, src/path.h:108-115
//   GetPathPss                    src/path.cpp:2588-2632
//
// Dropped on purpose (never read by the MLT loop): lensContrib / lensScore /
// lcJacobian / misWeight bookkeeping.  The degenerate-distance early-outs they guard are kept.
// Static scenes only: path.time is perturbed (it consumes RNG draws) but never moves geometry.
#pragma once
#include "light.h"
#include "rng.h"

namespace lmc {

struct SubpathContrib {          // src/path.h:12-21
    int camDepth, lightDepth;
    V2 screenPos;
    V3 contrib;
    float lsScore, ssScore;
};

struct alignas(16) SurfaceVertex {   // src/path.h:31-39; 48 B, moved as 3 x 16 B
    int tid;                     // shapeInst (triangle id in BVH order; -1 = none)
    V2 st;
    V2 bsdfRndParam;
    float bsdfDiscrete;
    float useAbsoluteParam;
    int dlLight, dlPrim;         // directLightInst
    V2 dlRndParam;               // directLightRndParam
    float rrWeight;
};

// Everything of a Path except the per-vertex arrays: 80 B, travels with the wavefront payload.
struct alignas(16) PathHead {
    float time;
    V2 screenPos;                // camVertex
    V2 lgtRndPos, lgtRndDir;     // lgtVertex
    int lgtLight, lgtPrim;
    int envLight, envPrim;       // envLightInst (envLight = -1: none)
    V3 lensVertexPos;
    int isSubpath, camDepth, lgtDepth;
    int nCam, nLgt;              // vector sizes
    int headPad;
};
template <int MAXD>
struct Path : PathHead {         // src/path.h:47-62
    SurfaceVertex cam[MAXD];
    SurfaceVertex lgt[MAXD];
};

LMC_HD void path_clear(PathHead &p) {
    p.nCam = 0; p.nLgt = 0; p.envLight = -1; p.envPrim = -1; p.isSubpath = 0;
}
template <int MAXD>
LMC_HD void path_copy(Path<MAXD> &d, const Path<MAXD> &s) {
    d.time = s.time; d.screenPos = s.screenPos; d.lgtRndPos = s.lgtRndPos; d.lgtRndDir = s.lgtRndDir;
    d.lgtLight = s.lgtLight; d.lgtPrim = s.lgtPrim; d.envLight = s.envLight; d.envPrim = s.envPrim;
    d.lensVertexPos = s.lensVertexPos; d.isSubpath = s.isSubpath; d.camDepth = s.camDepth;
    d.lgtDepth = s.lgtDepth; d.nCam = s.nCam; d.nLgt = s.nLgt;
    for (int i = 0; i < s.nCam; ++i) d.cam[i] = s.cam[i];
    for (int i = 0; i < s.nLgt; ++i) d.lgt[i] = s.lgt[i];
}
LMC_HD SurfaceVertex surface_vertex_zero() {
    SurfaceVertex v;
    v.tid = -1; v.st = mk2(0, 0); v.bsdfRndParam = mk2(0, 0); v.bsdfDiscrete = 0.0f;
    v.useAbsoluteParam = 0.0f; v.dlLight = -1; v.dlPrim = -1; v.dlRndParam = mk2(0, 0); v.rrWeight = 0.0f;
    return v;
}

LMC_HD int primary_param_size(int camDepth, int lightDepth) {
    const int l = camDepth + lightDepth - 1;
    return (l > 2 ? l : 2) * 2 + 1;
}
LMC_HD int path_dimension(const PathHead &p) { return primary_param_size(p.camDepth, p.lgtDepth) - 1; }

template <int MAXD>
LMC_HD void to_subpath(int camDepth, int lgtDepth, Path<MAXD> &path) {
    const int nc = camDepth - 1 > 0 ? camDepth - 1 : 0;
    const int nl = lgtDepth - 1 > 0 ? lgtDepth - 1 : 0;
    for (int i = path.nCam; i < nc; ++i) path.cam[i] = surface_vertex_zero();
    for (int i = path.nLgt; i < nl; ++i) path.lgt[i] = surface_vertex_zero();
    path.nCam = nc; path.nLgt = nl;
    if (lgtDepth != 0) path.envLight = -1;
    path.isSubpath = 1; path.camDepth = camDepth; path.lgtDepth = lgtDepth;
}

// GetPathPss: writes dim floats
template <int MAXD>
LMC_HD void get_path_pss(const Path<MAXD> &path, float *pss) {
    int k = 0;
    if (path.lgtDepth > 1) {
        pss[k++] = path.lgtRndPos.x; pss[k++] = path.lgtRndPos.y;
        pss[k++] = path.lgtRndDir.x; pss[k++] = path.lgtRndDir.y;
        for (int d = 0; d < path.nLgt; ++d) {
            if (d == path.nLgt - 1 && path.camDepth == 1) return;
            if (d == path.nLgt - 1) break;
            pss[k++] = path.lgt[d].bsdfRndParam.x; pss[k++] = path.lgt[d].bsdfRndParam.y;
        }
    }
    pss[k++] = path.screenPos.x; pss[k++] = path.screenPos.y;
    for (int d = 0; d < path.nCam; ++d) {
        if (d == path.nCam - 1) {
            if (path.lgtDepth == 1) { pss[k++] = path.cam[d].dlRndParam.x; pss[k++] = path.cam[d].dlRndParam.y; }
            return;
        }
        pss[k++] = path.cam[d].bsdfRndParam.x; pss[k++] = path.cam[d].bsdfRndParam.y;
    }
}

struct BidirPathState {
    Isect isect;
    V3 wi;
    float accMISWPrev, accMISWThis;
    V3 throughput;
    float ssJacobian;
    float lastBsdfPdf;
};

LMC_HD float mis(float pdf) { return square(pdf); }   // src/path.cpp:29-32

// ShadingNormalCorrection (src/path.cpp:34-54)
template <bool adjoint>
LMC_HD float shading_normal_correction(V3 wi, const Isect &isect, V3 wo) {
    const float cosWi = dot(isect.shadingNormal, wi);
    const float cosWo = dot(isect.shadingNormal, wo);
    const float wiDotGeoN = dot(isect.geomNormal, wi);
    const float woDotGeoN = dot(isect.geomNormal, wo);
    if (wiDotGeoN * cosWi <= 0.0f || woDotGeoN * cosWo <= 0.0f) return 0.0f;
    if (adjoint) return dm_abs((woDotGeoN * cosWi) / (wiDotGeoN * cosWo));
    return 1.0f;
}

LMC_HD void emit_from_camera(const Scene &sc, V2 screenPos, Ray &ray, float &minT, float &maxT,
                             BidirPathState &ps) {
    Ray centerRay; float cmin, cmax;
    camera_sample_primary(sc.cam, mk2(0.5f, 0.5f), centerRay, cmin, cmax);
    camera_sample_primary(sc.cam, screenPos, ray, minT, maxT);
    const V3 camDir = centerRay.dir;
    const V3 dir = ray.dir;
    const float cosAtCamera = dot(camDir, dir);
    const float imagePointToCameraDist = sc.cam.dist / cosAtCamera;
    const float imageToSolidAngleFactor = square(imagePointToCameraDist) / cosAtCamera;
    const float cameraPdfW = imageToSolidAngleFactor;
    const float screenPixelCount = (float)(sc.cam.width * sc.cam.height);
    ps.throughput = mk3s(1.0f);
    ps.accMISWPrev = mis(screenPixelCount / cameraPdfW);
    ps.accMISWThis = 0.0f;
    ps.ssJacobian = 1.0f;
}

LMC_HD void emit_from_light(const Scene &sc, float lightPickProb, PathHead &path, Ray &ray, BidirPathState &ps) {
    float cosLight, emissionPdf, directPdf;
    light_emit(sc, path.lgtLight, path.lgtRndPos, path.lgtRndDir, path.lgtPrim, ray, ps.throughput,
               cosLight, emissionPdf, directPdf);
    emissionPdf *= lightPickProb;
    directPdf *= lightPickProb;
    ps.throughput *= inverse(lightPickProb);
    ps.accMISWPrev = mis(directPdf / emissionPdf);
    if (!light_is_delta(sc.lights[path.lgtLight])) ps.accMISWThis = mis(cosLight / emissionPdf);
    else ps.accMISWThis = 0.0f;
    ps.ssJacobian = 1.0f;
}

// light == -1 encodes nullptr
LMC_HD void convert_mis(const Scene &sc, int depth, int light, const Ray &ray, BidirPathState &ps) {
    if (depth > 0 || light < 0 || light_is_finite(sc.lights[light])) {
        ps.accMISWPrev *= mis(distance_squared(ray.org, ps.isect.position));
    }
    const float invCosTheta = inverse(mis(dm_abs(dot(ray.dir, ps.isect.shadingNormal))));
    ps.accMISWPrev *= invCosTheta;
    ps.accMISWThis *= invCosTheta;
}

// Fixed-capacity contribution list (std::vector<SubpathContrib> in the reference)
template <int CAP>
struct ContribList {
    int n;
    SubpathContrib c[CAP];
    LMC_HD void clear() { n = 0; }
    LMC_HD void push(const SubpathContrib &x) { if (n < CAP) c[n++] = x; }
    // visibility of a connection segment: traced on the spot (see DeferredList in stages.h for
    // the wavefront variant, which queues the shadow ray and resolves the contribution later)
    LMC_HD bool occluded(const Scene &sc, const Ray &ray, float dist) { return scene_occluded(sc, ray, dist); }
    // ConnectVertex for one (camera vertex, light vertex) pair of GeneratePathBidir (src/path.cpp:1417-1431);
    // the wavefront's DeferredList may hand the pair to a dedicated kernel instead (stages.h)
    template <class LS>
    LMC_HD void connect(const Scene &sc, int camDepth, int lgtDepth, const LS *ls, const SurfaceVertex *lgtVerts,
                        const LS &cps, const SurfaceVertex &camVertex, V2 screenPos);
};

template <class CL>
LMC_HD_NOINLINE void connect_to_camera(const Scene &sc, int lgtDepth, const BidirPathState &ps,
                              const SurfaceVertex &lgtVertex, V3 prevPosition, CL &contribs) {
    Ray centerRay; float cmin, cmax;
    camera_sample_primary(sc.cam, mk2(0.5f, 0.5f), centerRay, cmin, cmax);
    const V3 camOrg = centerRay.org;
    const V3 camDir = centerRay.dir;
    V3 dirToCamera = camOrg - ps.isect.position;
    if (-dot(camDir, dirToCamera) <= 0.0f) return;
    V2 screenPos;
    if (!camera_project_point(sc.cam, ps.isect.position, screenPos)) return;
    const float distSq = length_squared(dirToCamera);
    const float dist = dm_sqrt(distSq);
    dirToCamera *= inverse(dist);
    // The visibility test has no side effects, so it is evaluated LAST (only for connections
    // that would contribute): same result as the reference's order (src/path.cpp:662-664).
    Ray sray; sray.org = ps.isect.position; sray.dir = dirToCamera;
    const int geom = sc.tris[lgtVertex.tid].geom;
    const BsdfParams bp = bsdf_params(sc, geom, lgtVertex.st);
    V3 bsdfContrib; float cosToCamera, bsdfPdf, bsdfRevPdf;
    bsdf_eval(true, bp, ps.wi, ps.isect.shadingNormal, dirToCamera, bsdfContrib, cosToCamera, bsdfPdf, bsdfRevPdf);
    if (is_zero(bsdfContrib)) return;
    const float factor = shading_normal_correction<true>(ps.wi, ps.isect, dirToCamera);
    if (factor <= 0.0f) return;
    bsdfContrib *= factor;
    const bool useAbsoluteParam = bsdf_roughness(bp) > sc.opt.roughnessThreshold;
    if (useAbsoluteParam && lgtDepth >= 1) {
        const float dsq = distance_squared(ps.isect.position, prevPosition);
        if (dsq <= 0.0f) { if (!contribs.occluded(sc, sray, dist)) contribs.clear(); return; }
    }
    const float cosAtCamera = -dot(camDir, dirToCamera);
    const float imagePointToCameraDist = sc.cam.dist / cosAtCamera;
    const float imageToSolidAngleFactor = square(imagePointToCameraDist) / cosAtCamera;
    const float imageToSurfaceFactor = imageToSolidAngleFactor * dm_abs(cosToCamera) / distSq;
    const float screenPixelCount = (float)(sc.cam.width * sc.cam.height);
    const float cameraPdf = imageToSurfaceFactor;
    const float wLight = mis(cameraPdf / screenPixelCount) * (ps.accMISWPrev + ps.accMISWThis * mis(bsdfRevPdf));
    const float misWeight = inverse(wLight + 1.0f);
    const float surfaceToImageFactor = cosToCamera / imageToSurfaceFactor;
    V3 contrib = misWeight * bsdfContrib / (screenPixelCount * surfaceToImageFactor);
    contrib = cmul(contrib, ps.throughput);
    const float score = luminance(contrib);
    if (score > 0.0f && !contribs.occluded(sc, sray, dist)) {
        SubpathContrib c;
        c.camDepth = 1; c.lightDepth = 2 + lgtDepth; c.screenPos = screenPos; c.contrib = contrib;
        c.lsScore = score; c.ssScore = score * ps.ssJacobian;
        contribs.push(c);
    }
}

// BSDFSampling<adjoint, perturb>.  `next` may alias `ps`.
template <bool adjoint, bool perturb>
LMC_HD_NOINLINE bool bsdf_sampling(const Scene &sc, const BidirPathState &ps, SurfaceVertex &sv,
                          BidirPathState &next, V3 &dir, V3 &bsdfContrib) {
    const int geom = sc.tris[sv.tid].geom;
    const BsdfParams bp = bsdf_params(sc, geom, sv.st);
    float cosWo, bsdfPdf, bsdfRevPdf;
    // The reference assigns nextPathState.ssJacobian only in the absolute-parametrisation
    // branches (src/path.cpp:790-798,822-827); after a sampled (glass) vertex the field keeps
    // whatever `next` held: the old value when next aliases ps, or the zero of a freshly
    // value-initialised BidirPathState() on the large-step light subpath (src/path.cpp:1289).
    bool setJacobian = false;
    float nextSsJacobian = 0.0f;
    sv.useAbsoluteParam = (bsdf_roughness(bp) > sc.opt.roughnessThreshold) ? 1.0f : 0.0f;
    if (!perturb || sv.useAbsoluteParam == 0.0f) {
        if (!bsdf_sample(adjoint, bp, ps.wi, ps.isect.shadingNormal, sv.bsdfRndParam, sv.bsdfDiscrete,
                         dir, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf)) return false;
        if (sv.useAbsoluteParam == 1.0f) {
            float jacobian;
            sv.bsdfRndParam = to_spherical_coord(dir, jacobian);
            jacobian *= bsdfPdf;
            nextSsJacobian = ps.ssJacobian * jacobian;
            setJacobian = true;
        }
    } else {
        float jacobian;
        dir = sample_sphere(sv.bsdfRndParam, jacobian);
        bsdf_eval(adjoint, bp, ps.wi, ps.isect.shadingNormal, dir, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf);
        if (is_zero(bsdfContrib) || bsdfPdf <= 0.0f) return false;
        bsdfContrib *= inverse(bsdfPdf);
        jacobian *= bsdfPdf;
        nextSsJacobian = ps.ssJacobian * jacobian;
        setJacobian = true;
    }
    const float factor = shading_normal_correction<adjoint>(ps.wi, ps.isect, dir);
    if (factor <= 0.0f) return false;
    bsdfContrib *= factor;
    const float accThis = mis(cosWo / bsdfPdf) * (ps.accMISWThis * mis(bsdfRevPdf) + ps.accMISWPrev);
    const V3 thr = cmul(ps.throughput, bsdfContrib);
    if (setJacobian) next.ssJacobian = nextSsJacobian;
    next.lastBsdfPdf = bsdfPdf;
    next.accMISWThis = accThis;
    next.accMISWPrev = mis(inverse(bsdfPdf));
    next.throughput = thr;
    return true;
}

template <class CL>
LMC_HD_NOINLINE void handle_hit_light(const Scene &sc, int camDepth, int light, bool hitSurface, const Ray &ray,
                             V2 screenPos, const BidirPathState &ps, PathHead &path, CL &contribs) {
    int lPrimID = -1;
    V3 emission; float directPdf, emissionPdf;
    light_emission(sc, light, ray.dir, ps.isect.shadingNormal, lPrimID, emission, directPdf, emissionPdf);
    if (emission.x + emission.y + emission.z > 0.0f) {
        V3 contrib = cmul(ps.throughput, emission);
        float misWeight = 1.0f;
        if (camDepth > 0) {
            const float lightPickProb = pick_light_prob(sc, light);
            directPdf *= lightPickProb;
            emissionPdf *= lightPickProb;
            const float wCamera = mis(directPdf) * ps.accMISWPrev + mis(emissionPdf) * ps.accMISWThis;
            misWeight = inverse(1.0f + wCamera);
            contrib *= misWeight;
        }
        const float score = luminance(contrib);
        if (score > 0.0f) {
            if (!hitSurface) { path.envLight = light; path.envPrim = lPrimID; }
            SubpathContrib c;
            c.camDepth = 2 + camDepth; c.lightDepth = 0; c.screenPos = screenPos; c.contrib = contrib;
            c.lsScore = score; c.ssScore = score * ps.ssJacobian;
            contribs.push(c);
        }
    }
}

template <class CL>
LMC_HD_NOINLINE void direct_lighting(const Scene &sc, int camDepth, const BidirPathState &ps, V2 screenPos,
                            float lightPickProb, SurfaceVertex &camVertex, CL &contribs) {
    const int light = camVertex.dlLight;
    V3 dirToLight, lightContrib; float dist, cosAtLight, directPdf, emissionPdf;
    if (!light_sample_direct(sc, light, ps.isect.position, camVertex.dlRndParam, camVertex.dlPrim,
                             dirToLight, dist, lightContrib, cosAtLight, directPdf, emissionPdf)) return;
    Ray sray; sray.org = ps.isect.position; sray.dir = dirToLight;   // visibility tested last (src/path.cpp:1003-1005)
    const int geom = sc.tris[camVertex.tid].geom;
    const BsdfParams bp = bsdf_params(sc, geom, camVertex.st);
    V3 bsdfContrib; float cosToLight, bsdfPdf, bsdfRevPdf;
    bsdf_eval(false, bp, ps.wi, ps.isect.shadingNormal, dirToLight, bsdfContrib, cosToLight, bsdfPdf, bsdfRevPdf);
    if (is_zero(bsdfContrib)) return;
    const float factor = shading_normal_correction<false>(ps.wi, ps.isect, dirToLight);
    if (factor <= 0.0f) return;
    bsdfContrib *= factor;
    V3 contrib = cmul(ps.throughput, bsdfContrib);
    contrib = cmul(contrib, lightContrib) * inverse(lightPickProb);
    const float wLight = light_is_delta(sc.lights[light]) ? 0.0f : mis(bsdfPdf / (lightPickProb * directPdf));
    const float wCamera = mis(emissionPdf * cosToLight / (directPdf * cosAtLight)) *
                          (ps.accMISWPrev + ps.accMISWThis * mis(bsdfRevPdf));
    const float misWeight = inverse(wLight + 1.0f + wCamera);
    contrib *= misWeight;
    const float score = luminance(contrib);
    if (score > 0.0f && !contribs.occluded(sc, sray, dist)) {
        SubpathContrib c;
        c.camDepth = 2 + camDepth; c.lightDepth = 1; c.screenPos = screenPos; c.contrib = contrib;
        c.lsScore = score; c.ssScore = score * ps.ssJacobian;
        contribs.push(c);
    }
}

template <class CL>
LMC_HD_NOINLINE void connect_vertex(const Scene &sc, int camDepth, int lgtDepth, const BidirPathState &lps,
                           const SurfaceVertex &lgtVertex, const BidirPathState &cps,
                           const SurfaceVertex &camVertex, V2 screenPos, CL &contribs) {
    V3 dirToLight = lps.isect.position - cps.isect.position;
    const float distSq = length_squared(dirToLight);
    const float dist = dm_sqrt(distSq);
    dirToLight *= inverse(dist);
    Ray sray; sray.org = cps.isect.position; sray.dir = dirToLight;   // visibility tested last (src/path.cpp:1116-1118)
    const BsdfParams cbp = bsdf_params(sc, sc.tris[camVertex.tid].geom, camVertex.st);
    V3 camBsdfFactor; float cosCamera, camBsdfPdf, camBsdfRevPdf;
    bsdf_eval(false, cbp, cps.wi, cps.isect.shadingNormal, dirToLight, camBsdfFactor, cosCamera, camBsdfPdf, camBsdfRevPdf);
    if (is_zero(camBsdfFactor)) return;
    const float camFactor = shading_normal_correction<false>(cps.wi, cps.isect, dirToLight);
    if (camFactor <= 0.0f) return;
    camBsdfFactor *= camFactor;
    const BsdfParams lbp = bsdf_params(sc, sc.tris[lgtVertex.tid].geom, lgtVertex.st);
    V3 lgtBsdfFactor; float cosLight, lgtBsdfPdf, lgtBsdfRevPdf;
    bsdf_eval(true, lbp, lps.wi, lps.isect.shadingNormal, -dirToLight, lgtBsdfFactor, cosLight, lgtBsdfPdf, lgtBsdfRevPdf);
    if (is_zero(lgtBsdfFactor)) return;
    const float lgtFactor = shading_normal_correction<true>(lps.wi, lps.isect, -dirToLight);
    if (lgtFactor <= 0.0f) return;
    lgtBsdfFactor *= lgtFactor;
    const float geometryTerm = inverse(distSq);
    const float camBsdfDirPdfA = camBsdfPdf * cosLight * geometryTerm;
    const float lgtBsdfDirPdfA = lgtBsdfPdf * cosCamera * geometryTerm;
    const float wLight = mis(camBsdfDirPdfA) * (lps.accMISWPrev + lps.accMISWThis * mis(lgtBsdfRevPdf));
    const float wCamera = mis(lgtBsdfDirPdfA) * (cps.accMISWPrev + cps.accMISWThis * mis(camBsdfRevPdf));
    const float misWeight = inverse(wLight + 1.0f + wCamera);
    const V3 throughput = cmul(lps.throughput, cps.throughput);
    V3 contrib = cmul(throughput, camBsdfFactor);
    contrib = cmul(contrib, lgtBsdfFactor) * geometryTerm;
    contrib *= misWeight;
    const float ssJacobian = lps.ssJacobian * cps.ssJacobian;
    const float score = luminance(contrib);
    if (score > 0.0f && !contribs.occluded(sc, sray, dist)) {
        SubpathContrib c;
        c.camDepth = 2 + camDepth; c.lightDepth = 2 + lgtDepth; c.screenPos = screenPos; c.contrib = contrib;
        c.lsScore = score; c.ssScore = score * ssJacobian;
        contribs.push(c);
    }
}

template <int CAP>
template <class LS>
LMC_HD void ContribList<CAP>::connect(const Scene &sc, int camDepth, int lgtDepth, const LS *ls, const SurfaceVertex *lgtVerts,
                                      const LS &cps, const SurfaceVertex &camVertex, V2 screenPos) {
    const SurfaceVertex lv = lgtVerts[lgtDepth];
    connect_vertex(sc, camDepth, lgtDepth, ls[lgtDepth], lv, cps, camVertex, screenPos, *this);
}

LMC_HD bool russian_roulette(int depth, V3 bsdfContrib, float &rrWeight, V3 &throughput, Rng &rng) {
    float rrProb = 1.0f;
    if (depth >= 3) rrProb = dm_min(max_coeff(bsdfContrib), 0.95f);
    if (rng_uniform(rng) > rrProb) return false;
    rrWeight = inverse(rrProb);
    throughput *= rrWeight;
    return true;
}

// hit light for a camera vertex: GetHitLight (src/path.cpp:104-120)
LMC_HD int get_hit_light(const Scene &sc, bool hitSurface, int tid) {
    if (!hitSurface) return sc.env.present ? sc.env.lightIndex : -1;
    if (tid >= 0) return sc.mats[sc.tris[tid].geom].areaLight;
    return -1;
}

// GeneratePathBidir(scene, (-1,-1), minDepth, maxDepth, path, contribs, rng)
template <int MAXD, class CL>
LMC_HD_NOINLINE void generate_path_bidir(const Scene &sc, int minDepth, int maxDepth, Path<MAXD> &path,
                                CL &contribs, Rng &rng) {
    path.time = rng_uniform(rng);
    BidirPathState lightStates[MAXD];
    int nLightStates = 1;
    float lightPickProb = 1.0f;
    // EmitFromLightInit
    path.lgtRndPos.x = rng_uniform(rng); path.lgtRndPos.y = rng_uniform(rng);
    path.lgtRndDir.x = rng_uniform(rng); path.lgtRndDir.y = rng_uniform(rng);
    path.lgtLight = pick_light(sc, rng_uniform(rng), lightPickProb);
    path.lgtPrim = light_sample_discrete(sc, path.lgtLight, rng_uniform(rng));
    Ray ray; float minT, maxT;
    emit_from_light(sc, lightPickProb, path, ray, lightStates[0]);
    minT = LMC_ISECT_EPS; maxT = dm_inf();
    for (int lgtDepth = 0;; lgtDepth++) {
        path.lgt[path.nLgt] = surface_vertex_zero();
        SurfaceVertex &sv = path.lgt[path.nLgt];
        path.nLgt++;
        BidirPathState &cur = lightStates[lgtDepth];
        const bool hitSurface = scene_intersect(sc, ray, minT, maxT, sv.tid, cur.isect, sv.st);
        if (!hitSurface) { nLightStates--; path.nLgt--; break; }
        sv.bsdfDiscrete = rng_uniform(rng);
        cur.wi = -ray.dir;
        convert_mis(sc, lgtDepth, path.lgtLight, ray, cur);
        if (lgtDepth + 2 >= minDepth) {
            connect_to_camera(sc, lgtDepth, cur, sv, ray.org, contribs);
        }
        if (maxDepth != -1 && lgtDepth + 2 >= maxDepth) break;
        nLightStates++;
        sv.bsdfRndParam.x = rng_uniform(rng); sv.bsdfRndParam.y = rng_uniform(rng);
        V3 bsdfContrib;
        lightStates[lgtDepth + 1].ssJacobian = 0.0f;   // BidirPathState() is zero-initialised
        if (!bsdf_sampling<true, false>(sc, cur, sv, lightStates[lgtDepth + 1], ray.dir, bsdfContrib)) {
            nLightStates--; break;
        }
        if (!russian_roulette(lgtDepth, bsdfContrib, sv.rrWeight, lightStates[lgtDepth + 1].throughput, rng)) {
            nLightStates--; break;
        }
        ray.org = cur.isect.position;
    }

    BidirPathState cps;
    path.screenPos.x = rng_uniform(rng); path.screenPos.y = rng_uniform(rng);   // EmitFromCameraInit, (-1,-1)
    emit_from_camera(sc, path.screenPos, ray, minT, maxT, cps);
    for (int camDepth = 0;; camDepth++) {
        path.cam[path.nCam] = surface_vertex_zero();
        SurfaceVertex &sv = path.cam[path.nCam];
        path.nCam++;
        const bool hitSurface = scene_intersect(sc, ray, minT, maxT, sv.tid, cps.isect, sv.st);
        cps.wi = -ray.dir;
        if (hitSurface) convert_mis(sc, camDepth, -1, ray, cps);
        if (camDepth + 1 >= minDepth) {
            const int light = get_hit_light(sc, hitSurface, sv.tid);
            if (light >= 0) {
                handle_hit_light(sc, camDepth, light, hitSurface, ray, path.screenPos, cps, path, contribs);
                return;
            }
        }
        if (!hitSurface || (maxDepth != -1 && camDepth + 1 >= maxDepth)) break;
        if (camDepth == 1) {
            path.lensVertexPos = cps.isect.position;
            const float distSq = distance_squared(cps.isect.position, ray.org);
            if (distSq <= 0.0f) { contribs.clear(); return; }
        }
        sv.bsdfDiscrete = rng_uniform(rng);
        if (camDepth + 2 >= minDepth) {
            float directLightPickProb = 1.0f;
            // DirectLightingInit
            sv.dlLight = pick_light(sc, rng_uniform(rng), directLightPickProb);
            sv.dlRndParam.x = rng_uniform(rng); sv.dlRndParam.y = rng_uniform(rng);
            sv.dlPrim = light_sample_discrete(sc, sv.dlLight, rng_uniform(rng));
            direct_lighting(sc, camDepth, cps, path.screenPos, directLightPickProb, sv, contribs);
        }
        int maxLgtDepth = nLightStates - 1;
        if (maxDepth != -1) {
            const int m = maxDepth - camDepth - 3;
            if (m < maxLgtDepth) maxLgtDepth = m;
        }
        for (int lgtDepth = 0; lgtDepth <= maxLgtDepth; lgtDepth++) {
            if (camDepth + lgtDepth + 3 >= minDepth) {
                contribs.connect(sc, camDepth, lgtDepth, lightStates, path.lgt, cps, sv, path.screenPos);
            }
        }
        sv.bsdfRndParam.x = rng_uniform(rng); sv.bsdfRndParam.y = rng_uniform(rng);
        V3 bsdfContrib;
        if (!bsdf_sampling<false, false>(sc, cps, sv, cps, ray.dir, bsdfContrib)) break;
        if (!russian_roulette(camDepth, bsdfContrib, sv.rrWeight, cps.throughput, rng)) break;
        ray.org = cps.isect.position;
        minT = LMC_ISECT_EPS; maxT = dm_inf();
    }
}

template <class OFF>
LMC_HD void perturb(float &value, const OFF &offset, int &offsetId) {
    value = modulo1(value + offset[offsetId++]);
}

// PerturbPathBidir(scene, offset, path, contribs, rng): contribs gets 0 or 1 entries.
template <int MAXD, class CL>
LMC_HD_NOINLINE void perturb_path_bidir(const Scene &sc, const float *offset, Path<MAXD> &path, CL &contribs, Rng &rng) {
    NormalDist normDist = normal_make(0.0f, sc.opt.discreteStdDev);
    int offsetId = 0;
    path.time = modulo1(path.time + normal_draw(normDist, rng));
    BidirPathState lps;
    if (path.lgtDepth > 1) {
        const float lightPickProb = pick_light_prob(sc, path.lgtLight);
        Ray ray;
        perturb(path.lgtRndPos.x, offset, offsetId);
        perturb(path.lgtRndPos.y, offset, offsetId);
        perturb(path.lgtRndDir.x, offset, offsetId);
        perturb(path.lgtRndDir.y, offset, offsetId);
        emit_from_light(sc, lightPickProb, path, ray, lps);
        const float minT = LMC_ISECT_EPS, maxT = dm_inf();
        for (int lgtDepth = 0; lgtDepth < path.nLgt; lgtDepth++) {
            SurfaceVertex &sv = path.lgt[lgtDepth];
            if (!scene_intersect(sc, ray, minT, maxT, sv.tid, lps.isect, sv.st)) return;
            lps.wi = -ray.dir;
            sv.bsdfDiscrete = modulo1(sv.bsdfDiscrete + normal_draw(normDist, rng));
            convert_mis(sc, lgtDepth, path.lgtLight, ray, lps);
            if (lgtDepth == path.nLgt - 1 && path.camDepth == 1) {
                connect_to_camera(sc, lgtDepth, lps, sv, ray.org, contribs);
                return;
            }
            if (lgtDepth == path.nLgt - 1) break;
            perturb(sv.bsdfRndParam.x, offset, offsetId);
            perturb(sv.bsdfRndParam.y, offset, offsetId);
            V3 bsdfContrib;
            if (!bsdf_sampling<true, true>(sc, lps, sv, lps, ray.dir, bsdfContrib)) return;
            lps.throughput *= sv.rrWeight;
            ray.org = lps.isect.position;
        }
    }
    perturb(path.screenPos.x, offset, offsetId);
    perturb(path.screenPos.y, offset, offsetId);
    Ray ray; float minT, maxT;
    BidirPathState cps;
    emit_from_camera(sc, path.screenPos, ray, minT, maxT, cps);
    for (int camDepth = 0; camDepth < path.nCam; camDepth++) {
        SurfaceVertex &sv = path.cam[camDepth];
        const bool hitSurface = scene_intersect(sc, ray, minT, maxT, sv.tid, cps.isect, sv.st);
        cps.wi = -ray.dir;
        if (hitSurface) convert_mis(sc, camDepth, -1, ray, cps);
        if (camDepth == path.nCam - 1 && path.lgtDepth == 0) {
            const int light = get_hit_light(sc, hitSurface, sv.tid);
            if (light >= 0) {
                handle_hit_light(sc, camDepth, light, hitSurface, ray, path.screenPos, cps, path, contribs);
            }
            return;
        }
        if (!hitSurface) return;
        sv.bsdfDiscrete = modulo1(sv.bsdfDiscrete + normal_draw(normDist, rng));
        if (camDepth == 1) {
            path.lensVertexPos = cps.isect.position;
            const float distSq = distance_squared(cps.isect.position, ray.org);
            if (distSq <= 0.0f) { contribs.clear(); return; }
        }
        if (camDepth == path.nCam - 1) {
            if (path.lgtDepth == 1) {
                const float directLightPickProb = pick_light_prob(sc, sv.dlLight);
                perturb(sv.dlRndParam.x, offset, offsetId);
                perturb(sv.dlRndParam.y, offset, offsetId);
                direct_lighting(sc, camDepth, cps, path.screenPos, directLightPickProb, sv, contribs);
            } else {
                connect_vertex(sc, camDepth, path.nLgt - 1, lps, path.lgt[path.nLgt - 1], cps, sv,
                               path.screenPos, contribs);
            }
            return;
        }
        perturb(sv.bsdfRndParam.x, offset, offsetId);
        perturb(sv.bsdfRndParam.y, offset, offsetId);
        V3 bsdfContrib;
        if (!bsdf_sampling<false, true>(sc, cps, sv, cps, ray.dir, bsdfContrib)) return;
        cps.throughput *= sv.rrWeight;
        ray.org = cps.isect.position;
        minT = LMC_ISECT_EPS; maxT = dm_inf();
    }
}

}  // namespace lmc
