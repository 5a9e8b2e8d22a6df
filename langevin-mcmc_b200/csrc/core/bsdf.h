// bsdf.h -- BSDF evaluate / sample (forward and adjoint) for the three reference BSDFs,
// enum-dispatched instead of virtual (SURVEY.md s8b "plugin surface kept").
//
// Reference: src/lambertian.cpp:15-90, src/phong.cpp:22-157, src/roughdielectric.cpp:22-330,
// src/microfacet.h:6-185, src/sampling.h:104-109, texture lookup src/bitmaptexture.h:73-97.
#pragma once
#include "scene.h"

namespace lmc {

#define LMC_PI 3.14159265358979323846f
#define LMC_INVPI (1.0f / LMC_PI)
#define LMC_TWOPI (2.0f * LMC_PI)
#define LMC_INVTWOPI (1.0f / LMC_TWOPI)
#define LMC_PIOVERTWO (0.5f * LMC_PI)
#define LMC_PIOVERFOUR (0.25f * LMC_PI)

// ---- textures ---------------------------------------------------------------------------
LMC_HD int imod(int a, int b) { const int r = a % b; return (r < 0) ? r + b : r; }

// Bilinear, periodic wrap, texel centres at (i + 0.5)/W; then the reference's
// fastpow(max(v, 0), gamma) (src/bitmaptexture.h:93-96).  OIIO's own filter is unpinned
// (SURVEY.md App. B#9); this is the restatement's definition.
LMC_HD_NOINLINE V3 texture_eval(const Scene &sc, int texId, V2 st) {
    const Texture &tx = sc.textures[texId];
    const float s = tx.sScale * st.x, t = tx.tScale * st.y;
    const float x = s * (float)tx.width - 0.5f, y = t * (float)tx.height - 0.5f;
    const float xf = dm_floor(x), yf = dm_floor(y);
    const float fx = x - xf, fy = y - yf;
    // floor values can be large; reduce in float before the int conversion
    const int x0 = imod((int)dm_fmod(xf, (float)tx.width), tx.width);
    const int y0 = imod((int)dm_fmod(yf, (float)tx.height), tx.height);
    const int x1 = (x0 + 1 == tx.width) ? 0 : x0 + 1;
    const int y1 = (y0 + 1 == tx.height) ? 0 : y0 + 1;
    const float *d = sc.texData + tx.offset;
    const V3 c00 = ld3(d + 3 * (y0 * tx.width + x0)), c10 = ld3(d + 3 * (y0 * tx.width + x1));
    const V3 c01 = ld3(d + 3 * (y1 * tx.width + x0)), c11 = ld3(d + 3 * (y1 * tx.width + x1));
    const V3 top = c00 * (1.0f - fx) + c10 * fx;
    const V3 bot = c01 * (1.0f - fx) + c11 * fx;
    const V3 v = top * (1.0f - fy) + bot * fy;
    return mk3(dm_fastpow(dm_max(v.x, 0.0f), tx.gamma), dm_fastpow(dm_max(v.y, 0.0f), tx.gamma),
               dm_fastpow(dm_max(v.z, 0.0f), tx.gamma));
}

LMC_HD V3 mat_kd(const Scene &sc, const Material &m, V2 st) {
    if (m.kdTex >= 0) return texture_eval(sc, m.kdTex, st);
    return ld3(m.Kd);
}

// Resolved BSDF parameters at a hit point == the reference's BSDF::Serialize(st, buffer)
// (src/lambertian.cpp:10-13, src/phong.cpp:14-20, src/roughdielectric.cpp:13-20).
struct BsdfParams {
    int type;
    int twoSided;
    V3 Kd, Ks, Kt;
    float exponent, KsWeight, eta, invEta, alpha;
};

LMC_HD BsdfParams bsdf_params(const Scene &sc, int geom, V2 st) {
    const Material &m = sc.mats[geom];
    BsdfParams p;
    p.type = m.type; p.twoSided = m.twoSided;
    p.Kd = (m.type == BSDF_ROUGHDIELECTRIC) ? mk3s(0.0f) : mat_kd(sc, m, st);
    p.Ks = (m.ksTex >= 0) ? texture_eval(sc, m.ksTex, st) : ld3(m.Ks);
    p.Kt = (m.ktTex >= 0) ? texture_eval(sc, m.ktTex, st) : ld3(m.Kt);
    p.exponent = (m.expTex >= 0) ? texture_eval(sc, m.expTex, st).x : m.exponent;
    p.KsWeight = m.KsWeight;
    p.eta = m.eta; p.invEta = m.invEta;
    p.alpha = (m.alphaTex >= 0) ? texture_eval(sc, m.alphaTex, st).x : m.alpha;
    return p;
}


// ---- sampling helpers ---------------------------------------------------------------------
// src/sampling.h:104-109 (ADEpsilon<Float>() == 0)
LMC_HD V3 sample_cos_hemisphere(V2 rnd) {
    const float phi = LMC_TWOPI * rnd.x;
    const float tmp = dm_sqrt(dm_max(1.0f - rnd.y, 0.0f));
    float sp, cp; dm_sincos(phi, sp, cp);
    return mk3(cp * tmp, sp * tmp, dm_sqrt(dm_max(rnd.y, 0.0f)));
}

// src/sampling.h:7-16
LMC_HD V3 sample_sphere(V2 coord, float &jacobian) {
    const float scaledTheta = LMC_TWOPI * coord.x;
    const float scaledPhi = LMC_PI * coord.y;
    float sinPhi, cosPhi; dm_sincos(scaledPhi, sinPhi, cosPhi);
    float st, ct; dm_sincos(scaledTheta, st, ct);
    jacobian = dm_abs(sinPhi) * LMC_TWOPI * LMC_PI;
    return mk3(sinPhi * ct, sinPhi * st, cosPhi);
}

// src/sampling.h:24-43
LMC_HD float patan2(float y, float x) {
    if (y == 0.0f && x == 0.0f) return 0.0f;
    float r = dm_atan2(y, x);
    if (r < 0.0f) r += LMC_TWOPI;
    return r;
}
LMC_HD V2 to_spherical_coord(V3 dir, float &jacobian) {
    const float theta = patan2(dir.y, dir.x) * LMC_INVTWOPI;
    float phi = dm_acos(dir.z);
    jacobian = dm_abs(dm_sin(phi)) * LMC_TWOPI * LMC_PI;
    phi *= LMC_INVPI;
    return mk2(theta, phi);
}

// ---- Lambertian (src/lambertian.cpp:15-90) ----------------------------------------------
LMC_HD void lambertian_eval(const BsdfParams &p, V3 wi, V3 normal, V3 wo,
                            V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    float cosWi = dot(normal, wi);
    V3 n = normal;
    if (p.twoSided && cosWi < 0.0f) { cosWi = -cosWi; n = -n; }
    cosWo = dot(n, wo);
    contrib = mk3s(0.0f);
    pdf = 0.0f; revPdf = 0.0f;   // reference leaves these unset; callers test contrib first
    if (cosWi < LMC_COS_EPS || cosWo < LMC_COS_EPS) return;
    const float fwdScalar = cosWo * LMC_INVPI;
    const float revScalar = cosWi * LMC_INVPI;
    contrib = fwdScalar * p.Kd;
    pdf = fwdScalar;
    revPdf = revScalar;
}

LMC_HD bool lambertian_sample(const BsdfParams &p, V3 wi, V3 normal, V2 rnd,
                              V3 &wo, V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    float cosWi = dot(wi, normal);
    V3 n = normal;
    if (dm_abs(cosWi) < LMC_COS_EPS) return false;
    if (cosWi < 0.0f) {
        if (p.twoSided) { cosWi = -cosWi; n = -n; } else return false;
    }
    V3 b0, b1;
    coordinate_system(n, b0, b1);
    const V3 r = sample_cos_hemisphere(rnd);
    wo = r.x * b0 + r.y * b1 + r.z * n;
    cosWo = r.z;
    pdf = r.z * LMC_INVPI;
    if (cosWo < LMC_COS_EPS) return false;
    revPdf = cosWi * LMC_INVPI;
    contrib = p.Kd;
    return true;
}

// ---- Phong (src/phong.cpp:22-157) ----------------------------------------------------------
LMC_HD void phong_eval(const BsdfParams &p, V3 wi, V3 normal, V3 wo,
                       V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    contrib = mk3s(0.0f); pdf = 0.0f; revPdf = 0.0f;
    float cosWi = dot(normal, wi);
    V3 n = normal;
    if (p.twoSided && cosWi < 0.0f) { cosWi = -cosWi; n = -n; }
    cosWo = dot(n, wo);
    if (cosWi <= LMC_COS_EPS || cosWo <= LMC_COS_EPS) return;
    if (p.KsWeight > 0.0f) {
        const float alpha = dm_max(dot(reflect(wi, n), wo), 0.0f);
        const float expo = p.exponent;
        const float weight = dm_pow(alpha, expo) * LMC_INVTWOPI;
        const float expoConst1 = expo + 1.0f;
        const float expoConst2 = expo + 2.0f;
        if (weight > 1e-10f) {
            contrib = p.Ks * (expoConst2 * weight);
            pdf = p.KsWeight * expoConst1 * weight;
            revPdf = pdf;
        }
    }
    if (p.KsWeight < 1.0f) {
        pdf += (1.0f - p.KsWeight) * cosWo * LMC_INVPI;
        revPdf += (1.0f - p.KsWeight) * cosWi * LMC_INVPI;
        contrib += p.Kd * LMC_INVPI;
    }
    contrib *= cosWo;
    if (max_coeff(contrib) < 1e-10f) contrib = mk3s(0.0f);
}

LMC_HD bool phong_sample(const BsdfParams &p, V3 wi, V3 normal, V2 rnd,
                         V3 &wo, V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    float cosWi = dot(wi, normal);
    if (dm_abs(cosWi) < LMC_COS_EPS) return false;
    V3 n_ = normal;
    if (cosWi < 0.0f) {
        if (p.twoSided) { cosWi = -cosWi; n_ = -n_; } else return false;
    }
    const float expo = p.exponent;
    const V3 R = reflect(wi, n_);
    float g; V3 n;
    const float uDiscrete = rnd.x;   // sic: lobe chosen from rndParam[0] (src/phong.cpp:97)
    float rndParam0;
    if (uDiscrete > p.KsWeight) {
        g = 1.0f; n = n_;
        rndParam0 = (uDiscrete - p.KsWeight) / (1.0f - p.KsWeight + 1e-10f);
    } else {
        g = expo; n = R;
        rndParam0 = uDiscrete / (p.KsWeight + 1e-10f);
    }
    const float power = 1.0f / (g + 1.0f);
    const float cosAlpha = dm_pow(rnd.y, power);
    const float sinAlpha = dm_sqrt(1.0f - square(cosAlpha));
    const float phi = LMC_TWOPI * rndParam0;
    float sp, cp; dm_sincos(phi, sp, cp);
    const V3 localDir = mk3(sinAlpha * cp, sinAlpha * sp, cosAlpha);
    V3 b0, b1;
    coordinate_system(n, b0, b1);
    wo = localDir.x * b0 + localDir.y * b1 + localDir.z * n;
    cosWo = dot(n_, wo);
    if (cosWo < LMC_COS_EPS) return false;
    contrib = mk3s(0.0f);
    pdf = 0.0f;
    revPdf = 0.0f;
    if (p.KsWeight > 0.0f) {
        const float alpha = dm_max(dot(R, wo), 0.0f);
        const float weight = dm_pow(alpha, expo) * LMC_INVTWOPI;
        const float expoConst1 = expo + 1.0f;
        const float expoConst2 = expo + 2.0f;
        if (weight > 1e-10f) {
            contrib = p.Ks * (expoConst2 * weight);
            pdf = p.KsWeight * expoConst1 * weight;
        }
        revPdf = pdf;
    }
    if (p.KsWeight < 1.0f) {
        contrib += p.Kd * LMC_INVPI;
        pdf += (1.0f - p.KsWeight) * cosWo * LMC_INVPI;
        revPdf += (1.0f - p.KsWeight) * cosWi * LMC_INVPI;
    }
    contrib *= cosWo;
    if (pdf < 1e-10f) return false;
    contrib *= inverse(pdf);
    return true;
}

// ---- microfacet helpers (src/microfacet.h) --------------------------------------------------
LMC_HD float beckmann_D(V3 localH, float alphaU, float alphaV) {
    const float cosTheta = localH.z, mu = localH.x, mv = localH.y;
    const float cosTheta2 = square(cosTheta);
    const float e = (square(mu) / square(alphaU) + square(mv) / square(alphaV)) / cosTheta2;
    return dm_exp(-e) / (LMC_PI * alphaU * alphaV * square(cosTheta2));
}
LMC_HD float beckmann_G1(float alpha, float cosTheta) {
    const float tanTheta = dm_sqrt(dm_abs(1.0f - square(cosTheta))) / cosTheta;
    if (tanTheta <= 0.0f) return 1.0f;
    const float a = 1.0f / (alpha * tanTheta);
    if (a >= 1.6f) return 1.0f;
    const float aSqr = a * a;
    return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
}
LMC_HD float beckmann_G(float alpha, float cosWi, float cosWo) {
    return beckmann_G1(alpha, cosWi) * beckmann_G1(alpha, cosWo);
}
LMC_HD float fresnel_dielectric_ext(float cosThetaI_, float &cosThetaT_, float eta, float invEta) {
    const float scale = (cosThetaI_ > 0.0f) ? invEta : eta;
    const float cosThetaTSqr = 1.0f - (1.0f - square(cosThetaI_)) * square(scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    const float cosThetaI = dm_abs(cosThetaI_);
    const float cosThetaT = dm_sqrt(cosThetaTSqr);
    const float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    const float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0.0f) ? -cosThetaT : cosThetaT;
    return 0.5f * (square(Rs) + square(Rp));
}
LMC_HD V3 sample_micronormal(V2 rnd, float alpha, float &pdfW) {
    const float phiM = LMC_TWOPI * rnd.y;
    float sinPhiM, cosPhiM; dm_sincos(phiM, sinPhiM, cosPhiM);
    const float alphaSqr = square(alpha);
    const float tanThetaMSqr = alphaSqr * (-dm_log(dm_max(1.0f - rnd.x, 1e-6f)));
    const float cosThetaM = 1.0f / dm_sqrt(1.0f + tanThetaMSqr);
    const float cosThetaMSqr = square(cosThetaM);
    pdfW = (1.0f - rnd.x) / (LMC_PI * alphaSqr * cosThetaM * cosThetaMSqr);
    const float sinThetaMSq = dm_max(1.0f - cosThetaMSqr, 0.0f);
    const float sinThetaM = dm_sqrt(sinThetaMSq);
    return mk3(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
}

// ---- RoughDielectric (src/roughdielectric.cpp:22-330) ------------------------------------
LMC_HD void roughdielectric_eval(bool adjoint, const BsdfParams &p, V3 wi, V3 normal, V3 wo,
                                 V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    const float cosWi = dot(wi, normal);
    contrib = mk3s(0.0f); cosWo = 0.0f; pdf = 0.0f; revPdf = 0.0f;
    if (dm_abs(cosWi) < LMC_COS_EPS) return;
    cosWo = dot(wo, normal);
    if (dm_abs(cosWo) < LMC_COS_EPS) return;
    const bool refl = cosWi * cosWo > 0.0f;
    const float eta_ = (cosWi > 0.0f) ? p.eta : p.invEta;
    const float revEta_ = (cosWo > 0.0f) ? p.eta : p.invEta;
    V3 H;
    if (refl) H = normalize(wi + wo); else H = normalize(wi + wo * eta_);
    if (dot(H, normal) < 0.0f) H = -H;
    const float cosHWi = dot(wi, H);
    const float cosHWo = dot(wo, H);
    if (dm_abs(cosHWi) < LMC_COS_EPS || dm_abs(cosHWo) < LMC_COS_EPS) return;
    if (cosHWi * cosWi <= 0.0f) return;
    if (cosHWo * cosWo <= 0.0f) return;
    V3 b0, b1;
    coordinate_system(normal, b0, b1);
    const V3 localH = mk3(dot(b0, H), dot(b1, H), dot(normal, H));
    const float alp = p.alpha;
    const float D = beckmann_D(localH, alp, alp);
    if (D <= 0.0f) return;
    const float revCosHWi = cosHWo;
    const float revCosHWo = cosHWi;
    float dummy;
    const float F = fresnel_dielectric_ext(cosHWi, dummy, p.eta, p.invEta);
    const float aCosWi = dm_abs(cosWi);
    const float aCosWo = dm_abs(cosWo);
    const float G = beckmann_G(alp, aCosWi, aCosWo);
    const float scaledAlpha = alp * (1.2f - 0.2f * dm_sqrt(aCosWi));
    const float scaledD = beckmann_D(localH, scaledAlpha, scaledAlpha);
    const float prob = localH.z * scaledD;
    if (prob < 1e-20f) { contrib = mk3s(0.0f); return; }
    const float revScaledAlpha = alp * (1.2f - 0.2f * dm_sqrt(aCosWo));
    const float revScaledD = beckmann_D(localH, revScaledAlpha, revScaledAlpha);
    const float revProb = localH.z * revScaledD;
    if (refl) {
        const float scalar = dm_abs(F * D * G / (4.0f * cosWi));
        contrib = p.Ks * scalar;
        pdf = dm_abs(prob * F / (4.0f * cosHWo));
        revPdf = dm_abs(revProb * F / (4.0f * revCosHWo));
    } else {
        const float sqrtDenom = cosHWi + eta_ * cosHWo;
        const float revSqrtDenom = revCosHWi + revEta_ * revCosHWo;
        const float factor = adjoint ? 1.0f : square(inverse(eta_));
        const float scalar = dm_abs(factor * ((1.0f - F) * D * G * square(eta_) * cosHWi * cosHWo) /
                                    (cosWi * square(sqrtDenom)));
        contrib = p.Kt * scalar;
        pdf = dm_abs(prob * (1.0f - F) * (square(eta_) * cosHWo) / (square(sqrtDenom)));
        revPdf = dm_abs(revProb * (1.0f - F) * (square(revEta_) * revCosHWo) / (square(revSqrtDenom)));
    }
}

LMC_HD bool roughdielectric_sample(bool adjoint, const BsdfParams &p, V3 wi, V3 normal, V2 rnd,
                                   float uDiscrete, V3 &wo, V3 &contrib, float &cosWo,
                                   float &pdf, float &revPdf) {
    const float cosWi = dot(wi, normal);
    if (dm_abs(cosWi) < LMC_COS_EPS) return false;
    const float alp = p.alpha;
    const float scaledAlp = alp * (1.2f - 0.2f * dm_sqrt(dm_abs(cosWi)));
    float mPdf;
    const V3 localH = sample_micronormal(rnd, scaledAlp, mPdf);
    pdf = mPdf;
    V3 b0, b1;
    coordinate_system(normal, b0, b1);
    const V3 H = localH.x * b0 + localH.y * b1 + localH.z * normal;
    const float cosHWi = dot(wi, H);
    if (dm_abs(cosHWi) < LMC_COS_EPS) return false;
    float cosThetaT = 0.0f;
    const float F = fresnel_dielectric_ext(cosHWi, cosThetaT, p.eta, p.invEta);
    const bool refl = uDiscrete <= F;
    V3 reflC;
    float cosHWo;
    if (refl) {
        wo = reflect(wi, H);
        if (F <= 0.0f || dot(normal, wo) * dot(normal, wi) <= 0.0f) return false;
        reflC = p.Ks;
        cosHWo = dot(wo, H);
        pdf = dm_abs(pdf * F / (4.0f * cosHWo));
        const float revCosHWo = cosHWi;
        const float rev_dwh_dwo = inverse(4.0f * revCosHWo);
        cosWo = dot(wo, normal);
        if (dm_abs(cosWo) < LMC_COS_EPS) return false;
        const float revScaledAlp = alp * (1.2f - 0.2f * dm_sqrt(dm_abs(cosWo)));
        const float revD = beckmann_D(localH, revScaledAlp, revScaledAlp);
        revPdf = dm_abs(F * revD * localH.z * rev_dwh_dwo);
    } else {
        wo = refract(wi, H, cosThetaT, p.eta, p.invEta);
        if (F >= 1.0f || cosThetaT == 0.0f || dot(normal, wo) * dot(normal, wi) >= 0.0f) return false;
        const float eta_ = (cosWi > 0.0f) ? p.eta : p.invEta;
        const float factor = adjoint ? 1.0f : square(inverse(eta_));
        reflC = p.Kt * factor;
        cosHWo = dot(wo, H);
        const float sqrtDenom = cosHWi + eta_ * cosHWo;
        const float dwh_dwo = (square(eta_) * cosHWo) / square(sqrtDenom);
        pdf = dm_abs(pdf * (1.0f - F) * dm_abs(dwh_dwo));
        cosWo = dot(wo, normal);
        if (dm_abs(cosWo) < LMC_COS_EPS) return false;
        const float revEta_ = (cosWo > 0.0f) ? p.eta : p.invEta;
        const float revCosHWi = cosHWo;
        const float revCosHWo = cosHWi;
        const float revSqrtDenom = revCosHWi + revEta_ * revCosHWo;
        const float rev_dwh_dwo = (square(revEta_) * revCosHWo) / square(revSqrtDenom);
        const float revScaledAlp = alp * (1.2f - 0.2f * dm_sqrt(dm_abs(cosWo)));
        const float revD = beckmann_D(localH, revScaledAlp, revScaledAlp);
        revPdf = dm_abs((1.0f - F) * revD * localH.z * rev_dwh_dwo);
    }
    if (dm_abs(cosHWo) < LMC_COS_EPS) return false;
    if (pdf < 1e-20f) return false;
    if (cosHWi * cosWi <= 0.0f) return false;
    if (cosHWo * cosWo <= 0.0f) return false;
    const float aCosWi = dm_abs(cosWi);
    const float aCosWo = dm_abs(cosWo);
    const float D = beckmann_D(localH, alp, alp);
    const float G = beckmann_G(alp, aCosWi, aCosWo);
    const float numerator = D * G * cosHWi;
    const float denominator = mPdf * aCosWi;
    contrib = reflC * dm_abs(numerator / denominator);
    return true;
}

// ---- the BSDF table (replaces the virtual calls of struct BSDF, src/bsdf.h:10-68) -----------------------------
// A device kernel cannot call a virtual per mutation, so the reference's plugin surface becomes an enum-keyed table:
// one row per BSDFType (include/lmc/bsdf.h mirrors the enum) naming the row's entry points, all with the SAME
// signatures:
//   <name>_evaluate (bool adjoint, const BsdfParams &, wi, normal, wo,                 contrib, cosWo, pdf, revPdf)
//   <name>_sample   (bool adjoint, const BsdfParams &, wi, normal, rnd, uDiscrete, wo,  contrib, cosWo, pdf, revPdf)
//   <name>_roughness(const BsdfParams &)
// The dispatchers below are generated from the table.  Adding a BSDF = a new enum value in core/scene.h, its three
// entry points, one row here, its parameters in BsdfParams / serialize_bsdf (the 10-float record of App. A.4), and its
// differentiable twin as a row of BSDF_TABLE in tools/adgen/pathfn.py (the adjoint is then generated, not written).
#define LMC_BSDF_TABLE(ROW) \
    ROW(BSDF_LAMBERTIAN, lambertian) \
    ROW(BSDF_PHONG, phong) \
    ROW(BSDF_ROUGHDIELECTRIC, roughdielectric)

LMC_HD void lambertian_evaluate(bool, const BsdfParams &p, V3 wi, V3 n, V3 wo, V3 &c, float &cw, float &pdf, float &rp) { lambertian_eval(p, wi, n, wo, c, cw, pdf, rp); }
LMC_HD bool lambertian_sample(bool, const BsdfParams &p, V3 wi, V3 n, V2 rnd, float, V3 &wo, V3 &c, float &cw, float &pdf, float &rp) { return lambertian_sample(p, wi, n, rnd, wo, c, cw, pdf, rp); }
LMC_HD float lambertian_roughness(const BsdfParams &) { return 1.0f; }                 // src/lambertian.h:37-39
LMC_HD void phong_evaluate(bool, const BsdfParams &p, V3 wi, V3 n, V3 wo, V3 &c, float &cw, float &pdf, float &rp) { phong_eval(p, wi, n, wo, c, cw, pdf, rp); }
LMC_HD bool phong_sample(bool, const BsdfParams &p, V3 wi, V3 n, V2 rnd, float, V3 &wo, V3 &c, float &cw, float &pdf, float &rp) { return phong_sample(p, wi, n, rnd, wo, c, cw, pdf, rp); }
LMC_HD float phong_roughness(const BsdfParams &) { return 1.0f; }                      // src/phong.cpp:155-157
LMC_HD void roughdielectric_evaluate(bool adjoint, const BsdfParams &p, V3 wi, V3 n, V3 wo, V3 &c, float &cw, float &pdf, float &rp) { roughdielectric_eval(adjoint, p, wi, n, wo, c, cw, pdf, rp); }
LMC_HD float roughdielectric_roughness(const BsdfParams &p) { return p.alpha; }        // src/roughdielectric.h:61-63

// BSDF::Evaluate / EvaluateAdjoint
LMC_HD_NOINLINE void bsdf_eval(bool adjoint, const BsdfParams &p, V3 wi, V3 normal, V3 wo,
                      V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    switch (p.type) {
#define LMC_ROW(ID, name) case ID: name##_evaluate(adjoint, p, wi, normal, wo, contrib, cosWo, pdf, revPdf); return;
        LMC_BSDF_TABLE(LMC_ROW)
#undef LMC_ROW
    }
    contrib = mk3s(0.0f); cosWo = 0.0f; pdf = 0.0f; revPdf = 0.0f;
}
// BSDF::Sample / SampleAdjoint
LMC_HD_NOINLINE bool bsdf_sample(bool adjoint, const BsdfParams &p, V3 wi, V3 normal, V2 rnd, float uDiscrete,
                        V3 &wo, V3 &contrib, float &cosWo, float &pdf, float &revPdf) {
    switch (p.type) {
#define LMC_ROW(ID, name) case ID: return name##_sample(adjoint, p, wi, normal, rnd, uDiscrete, wo, contrib, cosWo, pdf, revPdf);
        LMC_BSDF_TABLE(LMC_ROW)
#undef LMC_ROW
    }
    return false;
}
// BSDF::Roughness
LMC_HD float bsdf_roughness(const BsdfParams &p) {
    switch (p.type) {
#define LMC_ROW(ID, name) case ID: return name##_roughness(p);
        LMC_BSDF_TABLE(LMC_ROW)
#undef LMC_ROW
    }
    return 1.0f;
}

}  // namespace lmc
