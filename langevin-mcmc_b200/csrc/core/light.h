// light.h -- light sampling / emission (enum-dispatched) and the pinhole camera.
//
// Reference: src/envlight.cpp:24-248 (SampleDirection, SampleDirect, Emission, Emit),
// src/arealight.cpp:28-104, src/pointlight.cpp:22-72, src/trianglemesh.cpp:313-365 (Sample),
// src/distribution.h:8-63 (PiecewiseConstant1D), src/scene.cpp:151-158 (PickLight[Prob]),
// src/camera.cpp:38-51,67-84 (SamplePrimary, ProjectPoint), src/utils.h:261-267 (Tent),
// src/sampling.h:45-66 (SampleConcentricDisc).
#pragma once
#include "bsdf.h"
#include "bvh.h"

namespace lmc {

// first index i in [0, n) with a[i] >= u, else n   (std::lower_bound)
LMC_HD int lower_bound_f(const float *a, int n, float u) {
    int lo = 0, len = n;
    while (len > 0) {
        const int half = len >> 1;
        if (a[lo + half] < u) { lo = lo + half + 1; len = len - half - 1; } else { len = half; }
    }
    return lo;
}
// first index i in [0, n) with a[i] > u, else n   (std::upper_bound)
LMC_HD int upper_bound_f(const float *a, int n, float u) {
    int lo = 0, len = n;
    while (len > 0) {
        const int half = len >> 1;
        if (!(u < a[lo + half])) { lo = lo + half + 1; len = len - half - 1; } else { len = half; }
    }
    return lo;
}

// PiecewiseConstant1D::SampleDiscrete over a cdf of count+1 entries (src/distribution.h:44-50).
// pdf is returned through the stored cdf differences by the callers that need it.
LMC_HD int cdf_sample_discrete(const float *cdf, int count, float u) {
    const int p = upper_bound_f(cdf, count + 1, u);
    return dm_clampi(p - 1, 0, count - 1);
}

// PickLight (src/scene.cpp:151-154): prob = func[i] / (funcInt * count) = w_i / sum(w)
LMC_HD int pick_light(const Scene &sc, float u, float &prob) {
    const int id = cdf_sample_discrete(sc.lightPickCdf, sc.numLights, u);
    prob = sc.lights[id].samplingWeight / sc.lightWeightSum;
    return id;
}
LMC_HD float pick_light_prob(const Scene &sc, int lightId) {
    return sc.lights[lightId].samplingWeight / sc.lightWeightSum;
}

// Light::SampleDiscrete (src/light.h:21-23, src/arealight.cpp:24-26)
LMC_HD int light_sample_discrete(const Scene &sc, int lightId, float u) {
    const Light &l = sc.lights[lightId];
    if (l.type == LIGHT_AREA) return cdf_sample_discrete(sc.lightCdf + l.primCdfOffset, l.numPrims, u);
    return -1;
}

LMC_HD float tent(float s) {
    if (s < 0.5f) return 1.0f - dm_sqrt(2.0f * s);
    return dm_sqrt(2.0f * (s - 0.5f)) - 1.0f;
}

LMC_HD V2 sample_concentric_disc(V2 rnd) {
    const float r1 = 2.0f * rnd.x - 1.0f;
    const float r2 = 2.0f * rnd.y - 1.0f;
    float phi, r;
    if (r1 == 0.0f || r2 == 0.0f) { r = 0.0f; phi = 0.0f; }
    else if (square(r1) > square(r2)) { r = r1; phi = LMC_PIOVERFOUR * (r2 / r1); }
    else { r = r2; phi = LMC_PIOVERTWO - (r1 / r2) * LMC_PIOVERFOUR; }
    float sp, cp; dm_sincos(phi, sp, cp);
    return mk2(r * cp, r * sp);
}

// Image3::At with the linear index clamped into the image: the reference reads past the row /
// past the buffer at col == W-1 / row == H-1 (src/envlight.cpp:162-163).
LMC_HD V3 env_at_linear(const EnvMap &e, int x, int y) {
    int idx = y * e.width + x;
    const int last = e.width * e.height - 1;
    if (idx > last) idx = last;
    if (idx < 0) idx = 0;
    return ld3(e.image + 3 * idx);
}
LMC_HD V3 env_rep_at(const EnvMap &e, int x, int y) {
    return ld3(e.image + 3 * (imod(y, e.height) * e.width + imod(x, e.width)));
}

// SampleDirection (src/envlight.cpp:121-170)
LMC_HD void env_sample_direction(const EnvMap &e, V2 rnd, int &lPrimID, V3 &dirToLight, V3 &value, float &pdf) {
    float u0 = rnd.x, u1 = rnd.y;
    int row, col;
    {
        const int p = lower_bound_f(e.cdfRows, e.height + 1, u1);
        row = dm_clampi(p - 1, 0, e.height - 1);
        u1 = (u1 - e.cdfRows[row]) / (e.cdfRows[row + 1] - e.cdfRows[row]);
    }
    {
        const float *cdf = e.cdfCols + row * (e.width + 1);
        const int p = lower_bound_f(cdf, e.width + 1, u0);
        col = dm_clampi(p - 1, 0, e.width - 1);
        u0 = (u0 - cdf[col]) / (cdf[col + 1] - cdf[col]);
    }
    lPrimID = row * e.width + col;
    const float tx = tent(u0), ty = tent(u1);
    const float plx = (float)col + tx, ply = (float)row + ty;
    const float phi = (plx + 0.5f) * e.pixelSize[0];
    const float theta = (ply + 0.5f) * e.pixelSize[1];
    float sinPhi, cosPhi, sinTheta, cosTheta;
    dm_sincos(phi, sinPhi, cosPhi);
    dm_sincos(theta, sinTheta, cosTheta);
    dirToLight = xform_vector(e.toWorld, mk3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta));
    const float dx1 = tx, dx2 = 1.0f - tx, dy1 = ty, dy2 = 1.0f - ty;
    const V3 value1 = env_at_linear(e, col, row) * dx2 * dy2 + env_at_linear(e, col + 1, row) * dx1 * dy2;
    const V3 value2 = env_at_linear(e, col, row + 1) * dx2 * dy1 + env_at_linear(e, col + 1, row + 1) * dx1 * dy1;
    value = value1 + value2;
    const float rowWeight0 = e.rowWeights[dm_clampi(row, 0, e.height - 1)];
    const float rowWeight1 = e.rowWeights[dm_clampi(row + 1, 0, e.height - 1)];
    pdf = (luminance(value1) * rowWeight0 + luminance(value2) * rowWeight1) * e.normalization /
          dm_max(dm_abs(sinTheta), 1e-7f);
}

// TriangleMesh::Sample -> SampleDirect<Float> (src/trianglemesh.cpp:313-365)
LMC_HD void tri_sample(const Scene &sc, int tid, V2 rnd, V3 &pos, V3 &normal) {
    const TriGeom &tg = sc.tris[tid];
    const TriShade &ts = sc.shade[tid];
    const float a = dm_sqrt(1.0f - rnd.x);
    const float b1 = 1.0f - a;
    const float b2 = a * rnd.y;
    pos = ld3(tg.p0) + (ld3(tg.e1) * b1) + (ld3(tg.e2) * b2);
    normal = normalize(ld3(ts.n0) * (1.0f - b1 - b2) + ld3(ts.n1) * b1 + ld3(ts.n2) * b2);
}

LMC_HD int light_prim_tid(const Scene &sc, const Light &l, int lPrimID) {
    return sc.lightPrimTid[l.primTidOffset + lPrimID];
}

// Light::SampleDirect
LMC_HD_NOINLINE bool light_sample_direct(const Scene &sc, int lightId, V3 pos, V2 rnd, int &lPrimID,
                                V3 &dirToLight, float &dist, V3 &contrib, float &cosAtLight,
                                float &directPdf, float &emissionPdf) {
    const Light &l = sc.lights[lightId];
    if (l.type == LIGHT_ENV) {
        V3 value;
        env_sample_direction(sc.env, rnd, lPrimID, dirToLight, value, directPdf);
        dist = dm_inf();
        contrib = value * inverse(directPdf);
        cosAtLight = 1.0f;
        const float positionPdf = LMC_INVPI / square(sc.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
        return true;
    } else if (l.type == LIGHT_AREA) {
        V3 posOnLight, normalOnLight;
        const float shapePdf = l.invTotalArea;
        tri_sample(sc, light_prim_tid(sc, l, lPrimID), rnd, posOnLight, normalOnLight);
        dirToLight = posOnLight - pos;
        const float distSq = length_squared(dirToLight);
        dist = dm_sqrt(distSq);
        dirToLight = dirToLight / dist;
        cosAtLight = -dot(dirToLight, normalOnLight);
        if (cosAtLight > LMC_COS_EPS) {
            contrib = (cosAtLight / (distSq * shapePdf)) * ld3(l.emission);
            directPdf = shapePdf * distSq / cosAtLight;
            emissionPdf = shapePdf * cosAtLight * LMC_INVPI;
            return true;
        }
        return false;
    } else {
        dirToLight = ld3(l.pos) - pos;
        const float distSq = length_squared(dirToLight);
        directPdf = distSq;
        dist = dm_sqrt(distSq);
        dirToLight = dirToLight / dist;
        contrib = ld3(l.emission) * inverse(distSq);
        emissionPdf = 1.0f / (4.0f * LMC_PI);
        cosAtLight = 1.0f;
        lPrimID = 0;
        return true;
    }
}

// Light::Emission (env: src/envlight.cpp:195-226; area: src/arealight.cpp:62-79)
LMC_HD_NOINLINE void light_emission(const Scene &sc, int lightId, V3 dirToLight, V3 normalOnLight, int &lPrimID,
                           V3 &emission, float &directPdf, float &emissionPdf) {
    const Light &l = sc.lights[lightId];
    if (l.type == LIGHT_ENV) {
        const EnvMap &e = sc.env;
        const V3 d = xform_vector(e.toLight, dirToLight);
        const float uvx = dm_atan2(d.x, -d.z) * LMC_INVTWOPI * (float)e.width - 0.5f;
        const float uvy = dm_acos(d.y) * LMC_INVPI * (float)e.height - 0.5f;
        const int col = (int)dm_floor(uvx);
        const int row = (int)dm_floor(uvy);
        lPrimID = imod(row, e.height) * e.width + imod(col, e.width);
        const float dx1 = uvx - (float)col, dx2 = 1.0f - dx1, dy1 = uvy - (float)row, dy2 = 1.0f - dy1;
        const V3 value1 = env_rep_at(e, col, row) * dx2 * dy2 + env_rep_at(e, col + 1, row) * dx1 * dy2;
        const V3 value2 = env_rep_at(e, col, row + 1) * dx2 * dy1 + env_rep_at(e, col + 1, row + 1) * dx1 * dy1;
        emission = value1 + value2;
        const float sinTheta = dm_sqrt(1.0f - square(d.y));
        const float rowWeight0 = e.rowWeights[dm_clampi(row, 0, e.height - 1)];
        const float rowWeight1 = e.rowWeights[dm_clampi(row + 1, 0, e.height - 1)];
        directPdf = (luminance(value1) * rowWeight0 + luminance(value2) * rowWeight1) * e.normalization /
                    dm_max(dm_abs(sinTheta), 1e-7f);
        const float positionPdf = LMC_INVPI / square(sc.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
    } else if (l.type == LIGHT_AREA) {
        const float cosAtLight = -dot(normalOnLight, dirToLight);
        if (cosAtLight > 0.0f) {
            emission = ld3(l.emission);
            directPdf = l.invTotalArea;
            emissionPdf = cosAtLight * directPdf * LMC_INVPI;
        } else {
            emission = mk3s(0.0f); directPdf = 0.0f; emissionPdf = 0.0f;
        }
    } else {
        emission = mk3s(0.0f); directPdf = 0.0f; emissionPdf = 0.0f;
    }
}

// Light::Emit (env: src/envlight.cpp:228-248; area: src/arealight.cpp:81-104; point: src/pointlight.cpp:58-72)
LMC_HD_NOINLINE void light_emit(const Scene &sc, int lightId, V2 rndPos, V2 rndDir, int &lPrimID, Ray &ray,
                       V3 &emission, float &cosAtLight, float &emissionPdf, float &directPdf) {
    const Light &l = sc.lights[lightId];
    if (l.type == LIGHT_ENV) {
        env_sample_direction(sc.env, rndDir, lPrimID, ray.dir, emission, directPdf);
        ray.dir = -ray.dir;
        const V2 offset = sample_concentric_disc(rndPos);
        V3 b0, b1;
        coordinate_system(ray.dir, b0, b1);
        const V3 perpOffset = offset.x * b0 + offset.y * b1;
        ray.org = ld3(sc.bsphereCenter) + (perpOffset - ray.dir) * sc.bsphereRadius;
        cosAtLight = 1.0f;
        const float positionPdf = LMC_INVPI / square(sc.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
    } else if (l.type == LIGHT_AREA) {
        V3 normal;
        const float shapePdf = l.invTotalArea;
        tri_sample(sc, light_prim_tid(sc, l, lPrimID), rndPos, ray.org, normal);
        const V3 d = sample_cos_hemisphere(rndDir);
        V3 b0, b1;
        coordinate_system(normal, b0, b1);
        ray.dir = d.x * b0 + d.y * b1 + d.z * normal;
        emission = ld3(l.emission) * (LMC_PI / shapePdf);
        cosAtLight = d.z;
        emissionPdf = d.z * LMC_INVPI * shapePdf;
        directPdf = shapePdf;
    } else {
        ray.org = ld3(l.pos);
        float jac;
        ray.dir = sample_sphere(rndDir, jac);
        emission = ld3(l.emission);
        emissionPdf = 1.0f / (4.0f * LMC_PI);
        cosAtLight = 1.0f; directPdf = 1.0f;
    }
}

LMC_HD bool light_is_delta(const Light &l) { return l.type == LIGHT_POINT; }
LMC_HD bool light_is_finite(const Light &l) { return l.type != LIGHT_ENV; }

// ---- camera ---------------------------------------------------------------------------
// SamplePrimary (src/camera.cpp:38-51); static camera -> toWorld is a fixed matrix
LMC_HD void camera_sample_primary(const Camera &cam, V2 screenPos, Ray &ray, float &minT, float &maxT) {
    ray.org = xform_point(cam.sampleToCam, mk3(screenPos.x, screenPos.y, 0.0f));
    ray.dir = normalize(ray.org);
    const float invZ = inverse(ray.dir.z);
    ray.org = xform_point(cam.camToWorld, mk3s(0.0f));
    ray.dir = xform_vector(cam.camToWorld, ray.dir);
    minT = cam.nearClip * invZ;
    maxT = cam.farClip * invZ;
}
// ProjectPoint (src/camera.cpp:67-84)
LMC_HD bool camera_project_point(const Camera &cam, V3 p, V2 &screenPos) {
    const V3 camP = xform_point(cam.worldToCam, p);
    if (camP.z < cam.nearClip || camP.z > cam.farClip) return false;
    const V3 rasterP = xform_point(cam.camToSample, camP);
    if (rasterP.x < 0.0f || rasterP.x > 1.0f || rasterP.y < 0.0f || rasterP.y > 1.0f) return false;
    screenPos.x = rasterP.x;
    screenPos.y = rasterP.y;
    return true;
}

}  // namespace lmc
