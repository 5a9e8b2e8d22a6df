// direct.h -- the direct-lighting pre-pass of the MLT integrator: the
// unidirectional GeneratePath restricted to the depths DirectLighting() asks for, and the per-tile
// sample loop around it.
//
//   UniPathState / Init                src/path.cpp:72-90
//   HandleHitLight (unidirectional)    src/path.cpp:120-181
//   DirectLightingInit                 src/path.cpp:184-193
//   DirectLighting (unidirectional)    src/path.cpp:195-288
//   BSDFSampling (unidirectional)      src/path.cpp:291-385
//   GeneratePath                       src/path.cpp:406-527
//   DirectLighting(scene, buffer)      src/direct.cpp:4-54
//   MISWeight                          src/path.cpp:23-27
//
// Dropped on purpose: lensThroughput / lensScore / ssJacobian / lcJacobian bookkeeping (the pre-pass
// only splats `contrib`), light-coordinate sampling (useLightCoordinateSampling is rejected at load
// time) and motion blur (static scenes).  The RNG draws and their order are the reference's: one RNG
// per 16 x 16 tile seeded with the tile index (+ seedoffset), pixels of the tile and samples of a pixel
// in sequence -- which is also why one CUDA thread per tile is the natural (and bit-reproducible) mapping.
#pragma once
#include "mutation.h"

namespace lmc {

LMC_HD float mis_weight(float pdfA, float pdfB) {          // src/path.cpp:23-27
    const float ratioSq = square(pdfB / pdfA);
    return 1.0f / (1.0f + ratioSq);
}

struct UniPathState {
    Ray ray; float minT, maxT;
    V3 throughput;
    float lastBsdfPdf;
    Isect isect;
    V3 wi;
};

// HandleHitLight, unidirectional
template <class CL>
LMC_HD void uni_handle_hit_light(const Scene &sc, int camDepth, int light, bool hitSurface, const UniPathState &ps, V2 screenPos,
                                 CL &contribs) {
    int lPrimID = -1;
    V3 emission; float directPdf, emissionPdf;
    light_emission(sc, light, ps.ray.dir, ps.isect.shadingNormal, lPrimID, emission, directPdf, emissionPdf);
    if (emission.x + emission.y + emission.z > 0.0f) {
        if (hitSurface) {
            const float distSq = distance_squared(ps.ray.org, ps.isect.position);
            const float cosTheta = -dot(ps.ray.dir, ps.isect.shadingNormal);
            directPdf *= (distSq / cosTheta);
        }
        V3 contrib = cmul(ps.throughput, emission);
        if (camDepth > 0) {
            const float lightPickProb = pick_light_prob(sc, light);
            const float misWeight = mis_weight(ps.lastBsdfPdf, directPdf * lightPickProb);
            contrib *= misWeight;
        }
        const float score = luminance(contrib);
        if (score > 0.0f) {
            SubpathContrib c;
            c.camDepth = 2 + camDepth; c.lightDepth = 0; c.screenPos = screenPos; c.contrib = contrib;
            c.lsScore = score; c.ssScore = score;
            contribs.push(c);
        }
    }
}

// DirectLighting, unidirectional (doOcclusion = true)
template <class CL>
LMC_HD void uni_direct_lighting(const Scene &sc, int camDepth, const UniPathState &ps, V2 screenPos, float lightPickProb,
                                SurfaceVertex &sv, CL &contribs) {
    const int light = sv.dlLight;
    V3 dirToLight, lightContrib; float distToLight, cosAtLight, directPdf, emissionPdf;
    if (!light_sample_direct(sc, light, ps.isect.position, sv.dlRndParam, sv.dlPrim, dirToLight, distToLight, lightContrib,
                             cosAtLight, directPdf, emissionPdf)) return;
    Ray sray; sray.org = ps.isect.position; sray.dir = dirToLight;
    if (scene_occluded(sc, sray, distToLight)) return;
    const BsdfParams bp = bsdf_params(sc, sc.tris[sv.tid].geom, sv.st);
    V3 bsdfContrib; float cosWo, bsdfPdf, bsdfRevPdf;
    bsdf_eval(false, bp, ps.wi, ps.isect.shadingNormal, dirToLight, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf);
    if (is_zero(bsdfContrib)) return;
    V3 contrib = cmul(ps.throughput, bsdfContrib);
    contrib = cmul(contrib, lightContrib) * inverse(lightPickProb);
    if (!light_is_delta(sc.lights[light])) {
        const float misWeight = mis_weight(directPdf * lightPickProb, bsdfPdf);
        contrib *= misWeight;
    }
    const float score = luminance(contrib);
    if (score > 0.0f) {
        SubpathContrib c;
        c.camDepth = 2 + camDepth; c.lightDepth = 1; c.screenPos = screenPos; c.contrib = contrib;
        c.lsScore = score; c.ssScore = score;
        contribs.push(c);
    }
}

// BSDFSampling<perturb = false>, unidirectional
LMC_HD bool uni_bsdf_sampling(const Scene &sc, UniPathState &ps, SurfaceVertex &sv, V3 &bsdfContrib) {
    const BsdfParams bp = bsdf_params(sc, sc.tris[sv.tid].geom, sv.st);
    float cosWo, bsdfPdfRev;
    sv.useAbsoluteParam = (bsdf_roughness(bp) > sc.opt.roughnessThreshold) ? 1.0f : 0.0f;
    if (!bsdf_sample(false, bp, ps.wi, ps.isect.shadingNormal, sv.bsdfRndParam, sv.bsdfDiscrete, ps.ray.dir, bsdfContrib,
                     cosWo, ps.lastBsdfPdf, bsdfPdfRev)) return false;
    if (sv.useAbsoluteParam == 1.0f) {
        float jacobian;
        sv.bsdfRndParam = to_spherical_coord(ps.ray.dir, jacobian);
    }
    ps.throughput = cmul(ps.throughput, bsdfContrib);
    ps.ray.org = ps.isect.position;
    return true;
}

// GeneratePath(scene, (px, py), minDepth, maxDepth, path, contribs, rng)
template <class CL>
LMC_HD_NOINLINE void generate_path_uni(const Scene &sc, int px, int py, int minDepth, int maxDepth, CL &contribs, Rng &rng) {
    (void)rng_uniform(rng);                        // time
    V2 screenPos;
    screenPos.x = (px == -1) ? rng_uniform(rng) : (((float)px + rng_uniform(rng)) / (float)sc.cam.width);
    screenPos.y = (py == -1) ? rng_uniform(rng) : (((float)py + rng_uniform(rng)) / (float)sc.cam.height);
    UniPathState ps;
    ps.throughput = mk3s(1.0f);
    ps.lastBsdfPdf = 1.0f;
    ps.isect.position = mk3s(0.0f); ps.isect.shadingNormal = mk3s(0.0f); ps.isect.geomNormal = mk3s(0.0f);
    ps.wi = mk3s(0.0f);
    camera_sample_primary(sc.cam, screenPos, ps.ray, ps.minT, ps.maxT);
    for (int camDepth = 0;; camDepth++) {
        SurfaceVertex sv = surface_vertex_zero();
        const bool hitSurface = scene_intersect(sc, ps.ray, ps.minT, ps.maxT, sv.tid, ps.isect, sv.st);
        const int light = get_hit_light(sc, hitSurface, sv.tid);
        if (light >= 0) {
            if (camDepth + 1 >= minDepth) {
                uni_handle_hit_light(sc, camDepth, light, hitSurface, ps, screenPos, contribs);
                return;
            }
        }
        if (!hitSurface || (maxDepth != -1 && camDepth + 1 >= maxDepth)) break;
        sv.bsdfDiscrete = rng_uniform(rng);
        if (camDepth == 1) {
            const float distSq = distance_squared(ps.isect.position, ps.ray.org);
            if (distSq <= 0.0f) { contribs.clear(); return; }
        }
        ps.wi = -ps.ray.dir;
        if (camDepth + 2 >= minDepth) {
            float directLightPickProb = 1.0f;
            sv.dlLight = pick_light(sc, rng_uniform(rng), directLightPickProb);
            sv.dlRndParam.x = rng_uniform(rng); sv.dlRndParam.y = rng_uniform(rng);
            sv.dlPrim = light_sample_discrete(sc, sv.dlLight, rng_uniform(rng));
            uni_direct_lighting(sc, camDepth, ps, screenPos, directLightPickProb, sv, contribs);
        }
        sv.bsdfRndParam.x = rng_uniform(rng); sv.bsdfRndParam.y = rng_uniform(rng);
        V3 bsdfContrib;
        if (!uni_bsdf_sampling(sc, ps, sv, bsdfContrib)) break;
        if (!russian_roulette(camDepth, bsdfContrib, sv.rrWeight, ps.throughput, rng)) break;
        ps.minT = LMC_ISECT_EPS; ps.maxT = dm_inf();
    }
}

#define LMC_DIRECT_TILE 16
// One tile of DirectLighting(scene, buffer) (the body of the ParallelFor lambda, src/direct.cpp:21-47).
// `rng` must have been seeded with tileY * nXTiles + tileX + seedOffset.
template <class FILM>
LMC_HD void direct_lighting_tile(const Scene &sc, int tileX, int tileY, int directSpp, Rng &rng, FILM &film) {
    const int W = sc.cam.width, H = sc.cam.height;
    const int x0 = tileX * LMC_DIRECT_TILE, x1 = (x0 + LMC_DIRECT_TILE < W) ? x0 + LMC_DIRECT_TILE : W;
    const int y0 = tileY * LMC_DIRECT_TILE, y1 = (y0 + LMC_DIRECT_TILE < H) ? y0 + LMC_DIRECT_TILE : H;
    const int minDepth = sc.opt.minDepth < 2 ? sc.opt.minDepth : 2;
    const int maxDepth = sc.opt.maxDepth < 2 ? sc.opt.maxDepth : 2;
    for (int y = y0; y < y1; y++) {
        for (int x = x0; x < x1; x++) {
            for (int s = 0; s < directSpp; s++) {
                ContribList<4> contribs; contribs.clear();
                generate_path_uni(sc, x, y, minDepth, maxDepth, contribs, rng);
                for (int i = 0; i < contribs.n; i++) splat(film, W, H, contribs.c[i].screenPos, contribs.c[i].contrib);
            }
        }
    }
}
// DirectLighting() is a no-op for these options (src/direct.cpp:6-8)
LMC_HD bool direct_lighting_skipped(const Scene &sc) { return sc.opt.minDepth > 2 || sc.opt.maxDepth < 1; }

}  // namespace lmc
