// pathgrad.h -- log-luminance of a path and its gradient w.r.t. the primary sample space,
// evaluated from the reference's serialized path buffers.
//
// This replaces the chad-generated functions `evaluate_path_bidir_mala_<c>_<l>_static[_derv]`
// (spec: RegisterPathFuncBidirMALA, src/path.cpp:3664-3911; ABI src/path.h:121-125).  It follows
// the reference's *AD twin* of the path sampler, which differs from the tracer on purpose:
// no validity early-outs, no occlusion, epsilon guards, (f cos x Jacobian) instead of
// (f cos / pdf) at absolutely-parametrised vertices:
//   EmitFromLight/ConvertMIS*/ConnectToCamera/BSDFSampling/EmitFromCamera/HandleHitLight/
//   DirectLighting/ConnectVertex                     src/path.cpp:2799-3380
//   TriangleIntersect / SampleDirect (shape)          src/trianglemesh.cpp:81-105,313-327
//   Evaluate*/Sample* BSDF twins                      src/lambertian.cpp:95-151, src/phong.cpp:171-393,
//                                                     src/roughdielectric.cpp:332-528, src/microfacet.h
//   light twins                                       src/envlight.cpp:250-399, src/arealight.cpp:106-208,
//                                                     src/pointlight.cpp:74-116
//   SamplePrimary twin                                src/camera.cpp:53-66
//
// Differentiation is forward mode with N-wide dual numbers carried in registers (no tape, no
// generated code): one templated statement list serves the plain-float forward value and the
// dual-number gradient, N directions per sweep.  Derivative conventions are chad's
// (src/chad.h): d|x| = +1 for x >= 0, fmax picks the first argument on ties, branches are
// differentiated on the taken side.  The result is the TRUE gradient of the function above; the
// reference's reverse-mode code deviates from it on paths through RoughDielectric vertices
// (SURVEY.md App. B#13), which tests/ quantify against oracle/_ref.
#pragma once
#include "bsdf.h"

namespace lmc {

// ------------------------------------------------------------------------------------------
// dual numbers over a scalar S (float for first order; a dual itself for second order)
// ------------------------------------------------------------------------------------------
template <class S, int N>
struct DualT {
    S v;
    S d[N];
};
template <int N> using Dual = DualT<float, N>;

// value access / construction usable with T = float too
LMC_HD float ad_val(float a) { return a; }
template <class S, int N> LMC_HD float ad_val(const DualT<S, N> &a) { return ad_val(a.v); }
template <class T> struct ADTraits;
template <> struct ADTraits<float> { LMC_HD static float make(float c) { return c; } };
template <class S, int N> struct ADTraits<DualT<S, N>> {
    LMC_HD static DualT<S, N> make(float c) {
        DualT<S, N> r; r.v = ADTraits<S>::make(c);
        for (int i = 0; i < N; i++) r.d[i] = ADTraits<S>::make(0.0f);
        return r;
    }
};
template <class T> LMC_HD T ad_const(float c) { return ADTraits<T>::make(c); }

#define LMC_DT template <class S, int N> LMC_HD DualT<S, N>
// a * b + c: one fused multiply-add on floats (explicit, see vec.h), plain composition on duals
LMC_HD float ad_mad(float a, float b, float c) { return fmaf(a, b, c); }
template <class S, int N> LMC_HD DualT<S, N> operator+(const DualT<S, N> &a, const DualT<S, N> &b);
template <class S, int N> LMC_HD DualT<S, N> operator*(const DualT<S, N> &a, const DualT<S, N> &b);
template <class S, int N> LMC_HD DualT<S, N> ad_mad(const DualT<S, N> &a, const DualT<S, N> &b, const DualT<S, N> &c) { return a * b + c; }
LMC_DT operator+(const DualT<S, N> &a, const DualT<S, N> &b) { DualT<S, N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
LMC_DT operator-(const DualT<S, N> &a, const DualT<S, N> &b) { DualT<S, N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
LMC_DT operator-(const DualT<S, N> &a) { DualT<S, N> r; r.v = -a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
LMC_DT operator*(const DualT<S, N> &a, const DualT<S, N> &b) { DualT<S, N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = ad_mad(a.d[i], b.v, a.v * b.d[i]); return r; }
LMC_DT operator/(const DualT<S, N> &a, const DualT<S, N> &b) {
    DualT<S, N> r; const S ib = 1.0f / b.v; r.v = a.v * ib;
    for (int i = 0; i < N; i++) r.d[i] = ad_mad(-r.v, b.d[i], a.d[i]) * ib;
    return r;
}
LMC_DT operator+(const DualT<S, N> &a, float b) { DualT<S, N> r = a; r.v = a.v + b; return r; }
LMC_DT operator+(float b, const DualT<S, N> &a) { DualT<S, N> r = a; r.v = b + a.v; return r; }
LMC_DT operator-(const DualT<S, N> &a, float b) { DualT<S, N> r = a; r.v = a.v - b; return r; }
LMC_DT operator-(float b, const DualT<S, N> &a) { DualT<S, N> r; r.v = b - a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
LMC_DT operator*(const DualT<S, N> &a, float b) { DualT<S, N> r; r.v = a.v * b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; }
LMC_DT operator*(float b, const DualT<S, N> &a) { DualT<S, N> r; r.v = b * a.v; for (int i = 0; i < N; i++) r.d[i] = b * a.d[i]; return r; }
LMC_DT operator/(const DualT<S, N> &a, float b) { DualT<S, N> r; const float ib = 1.0f / b; r.v = a.v * ib; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * ib; return r; }
LMC_DT operator/(float a, const DualT<S, N> &b) {
    DualT<S, N> r; const S ib = 1.0f / b.v; r.v = a * ib; const S k = -(r.v * ib);
    for (int i = 0; i < N; i++) r.d[i] = k * b.d[i];
    return r;
}
template <class S, int N> LMC_HD DualT<S, N> &operator+=(DualT<S, N> &a, const DualT<S, N> &b) { a = a + b; return a; }
template <class S, int N> LMC_HD DualT<S, N> &operator*=(DualT<S, N> &a, const DualT<S, N> &b) { a = a * b; return a; }
template <class S, int N> LMC_HD DualT<S, N> &operator*=(DualT<S, N> &a, float b) { a = a * b; return a; }

// chain rule helper: value v, derivative parts scaled by k (both of the scalar type S)
LMC_DT dchain(const DualT<S, N> &a, const S &v, const S &k) { DualT<S, N> r; r.v = v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * k; return r; }

LMC_HD float ad_sqrt(float a) { return dm_sqrt(a); }
LMC_DT ad_sqrt(const DualT<S, N> &a) { const S s = ad_sqrt(a.v); return dchain(a, s, 0.5f / s); }
LMC_HD float ad_sin(float a) { return dm_sin(a); }
LMC_HD float ad_cos(float a) { return dm_cos(a); }
// sin and cos of the same argument with ONE evaluation of the underlying kernel
LMC_HD void ad_sincos(float a, float &s, float &c) { dm_sincos(a, s, c); }
template <class S, int N> LMC_HD void ad_sincos(const DualT<S, N> &a, DualT<S, N> &s, DualT<S, N> &c) {
    S sv, cv; ad_sincos(a.v, sv, cv);
    s = dchain(a, sv, cv); c = dchain(a, cv, -sv);
}
LMC_DT ad_sin(const DualT<S, N> &a) { DualT<S, N> s, c; ad_sincos(a, s, c); return s; }
LMC_DT ad_cos(const DualT<S, N> &a) { DualT<S, N> s, c; ad_sincos(a, s, c); return c; }
LMC_HD float ad_exp(float a) { return dm_exp(a); }
LMC_DT ad_exp(const DualT<S, N> &a) { const S e = ad_exp(a.v); return dchain(a, e, e); }
LMC_HD float ad_log(float a) { return dm_log(a); }
LMC_DT ad_log(const DualT<S, N> &a) { return dchain(a, ad_log(a.v), 1.0f / a.v); }
// pow(x, y) with a constant exponent, libm semantics for a negative base with an integral
// exponent (Phong's `pow(alpha, exponent)` has no max(alpha, 0) in the AD twin, src/phong.cpp:202).
LMC_HD float ad_pow_c(float x, float y) {
    if (x < 0.0f) {
        const float fy = dm_floor(y);
        if (fy != y) return dm_nan();
        const float m = dm_pow(-x, y);
        const float half = y * 0.5f;
        return (dm_floor(half) != half) ? -m : m;
    }
    return dm_pow(x, y);
}
LMC_HD float ad_pow(float x, float y) { return ad_pow_c(x, y); }
// d/dx = y * pow(x, y - 1)   (src/chad.h:726-728)
LMC_DT ad_pow(const DualT<S, N> &x, float y) { return dchain(x, ad_pow(x.v, y), y * ad_pow(x.v, y - 1.0f)); }
LMC_HD float ad_fabs(float a) { return (a >= 0.0f) ? a : -a; }
LMC_DT ad_fabs(const DualT<S, N> &a) { return (ad_val(a) >= 0.0f) ? a : -a; }
LMC_HD float ad_fmax(float a, float b) { return (a >= b) ? a : b; }
LMC_DT ad_fmax(const DualT<S, N> &a, float b) { return (ad_val(a) >= b) ? a : ADTraits<DualT<S, N>>::make(b); }
LMC_HD float ad_atan2(float y, float x) { return dm_atan2(y, x); }
LMC_DT ad_atan2(const DualT<S, N> &y, const DualT<S, N> &x) {
    const S invNorm = 1.0f / (x.v * x.v + y.v * y.v);
    DualT<S, N> r; r.v = ad_atan2(y.v, x.v);
    const S ky = x.v * invNorm, kx = -(y.v * invNorm);
    for (int i = 0; i < N; i++) r.d[i] = ad_mad(ky, y.d[i], kx * x.d[i]);
    return r;
}
LMC_HD float ad_acos(float a) { return dm_acos(a); }
LMC_DT ad_acos(const DualT<S, N> &a) { return dchain(a, ad_acos(a.v), -(1.0f / ad_sqrt(1.0f - a.v * a.v))); }
template <class T> LMC_HD T ad_square(const T &a) { return a * a; }
template <class T> LMC_HD T ad_inverse(const T &a) { return 1.0f / a; }
template <class T> LMC_HD T ad_mis(const T &a) { return a * a; }

// ------------------------------------------------------------------------------------------
// small vectors over T
// ------------------------------------------------------------------------------------------
template <class T> struct TV3 { T x, y, z; };
template <class T> LMC_HD TV3<T> tv3(const T &x, const T &y, const T &z) { TV3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> LMC_HD TV3<T> tv3c(V3 v) { return tv3(ad_const<T>(v.x), ad_const<T>(v.y), ad_const<T>(v.z)); }
template <class T> LMC_HD TV3<T> operator+(const TV3<T> &a, const TV3<T> &b) { return tv3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> LMC_HD TV3<T> operator-(const TV3<T> &a, const TV3<T> &b) { return tv3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> LMC_HD TV3<T> operator-(const TV3<T> &a) { return tv3<T>(-a.x, -a.y, -a.z); }
template <class T> LMC_HD TV3<T> operator*(const TV3<T> &a, const T &s) { return tv3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> LMC_HD TV3<T> operator*(const T &s, const TV3<T> &a) { return tv3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> LMC_HD TV3<T> tcmul(const TV3<T> &a, const TV3<T> &b) { return tv3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class T> LMC_HD T tdot(const TV3<T> &a, const TV3<T> &b) { return ad_mad(a.z, b.z, ad_mad(a.y, b.y, a.x * b.x)); }
template <class T> LMC_HD TV3<T> tcross(const TV3<T> &a, const TV3<T> &b) {
    return tv3<T>(ad_mad(a.y, b.z, -(a.z * b.y)), ad_mad(a.z, b.x, -(a.x * b.z)), ad_mad(a.x, b.y, -(a.y * b.x)));
}
template <class T> LMC_HD T tlength_squared(const TV3<T> &v) { return ad_mad(v.z, v.z, ad_mad(v.y, v.y, v.x * v.x)); }
template <class T> LMC_HD TV3<T> tnormalize(const TV3<T> &v) {
    const T il = ad_inverse(ad_sqrt(tlength_squared(v)));
    return v * il;
}
template <class T> LMC_HD T tdistance_squared(const TV3<T> &a, const TV3<T> &b) {
    const T dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return ad_mad(dz, dz, ad_mad(dy, dy, dx * dx));
}
template <class T> LMC_HD T tluminance(const TV3<T> &v) { return v.x * 0.212671f + v.y * 0.715160f + v.z * 0.072169f; }
// constant-vector (V3) mixed helpers
template <class T> LMC_HD TV3<T> tscale(V3 c, const T &s) { return tv3<T>(c.x * s, c.y * s, c.z * s); }
template <class T> LMC_HD TV3<T> tscalef(const TV3<T> &a, float s) { return tv3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> LMC_HD T tdotc(V3 c, const TV3<T> &a) { return c.x * a.x + c.y * a.y + c.z * a.z; }

template <class T> LMC_HD void tcoordinate_system(const TV3<T> &n, TV3<T> &b1, TV3<T> &b2) {
    if (ad_val(n.z) < (float)(-1.0 + 1e-6)) {
        b1 = tv3c<T>(mk3(0.0f, -1.0f, 0.0f));
        b2 = tv3c<T>(mk3(-1.0f, 0.0f, 0.0f));
        return;
    }
    const T a = 1.0f / (1.0f + n.z);
    const T b = -n.x * n.y * a;
    b1 = tv3<T>(1.0f - ad_square(n.x) * a, b, -n.x);
    b2 = tv3<T>(b, 1.0f - ad_square(n.y) * a, -n.y);
}
template <class T> LMC_HD TV3<T> treflect(const TV3<T> &wi, const TV3<T> &n) { return (2.0f * tdot(wi, n)) * n - wi; }
template <class T> LMC_HD TV3<T> trefract(const TV3<T> &wi, const TV3<T> &n, const T &cosThetaT, float eta, float invEta) {
    const float eta_ = (ad_val(cosThetaT) < 0.0f) ? invEta : eta;
    return n * (tdot(wi, n) * eta_ + cosThetaT) - tscalef(wi, eta_);
}
// 3x3 part of a constant row-major matrix applied to a T vector
template <class T> LMC_HD TV3<T> txform_vector(const M44 &t, const TV3<T> &v) {
    return tv3<T>(t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z,
                  t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
                  t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z);
}
template <class T> LMC_HD TV3<T> txform_point(const M44 &t, const TV3<T> &p) {
    const T x = t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3];
    const T y = t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3];
    const T z = t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3];
    const T w = t.m[3][0] * p.x + t.m[3][1] * p.y + t.m[3][2] * p.z + t.m[3][3];
    const T iw = ad_inverse(w);
    return tv3<T>(x * iw, y * iw, z * iw);
}

// ------------------------------------------------------------------------------------------
// serialized-buffer layout constants (src/trianglemesh.cpp:3-10, src/bsdf.cpp:7-11, src/light.cpp:7-10)
// ------------------------------------------------------------------------------------------
#define LMC_SER_SHAPE 46
#define LMC_SER_BSDF 10
#define LMC_SER_LIGHT 56
#define LMC_SER_SCENE 38

// AnimatedTransform (15 floats) -> static matrix: Translate(translate[0]) * ToMatrix4x4(rotate[0])
// (src/animatedtransform.cpp:33-38,66-68, src/quaternion.h:13-38)
LMC_HD M44 ser_static_matrix(const float *a15) {
    const float *t = a15 + 1, *q = a15 + 7;
    const float xx = q[0] * q[0], yy = q[1] * q[1], zz = q[2] * q[2];
    const float xy = q[0] * q[1], xz = q[0] * q[2], yz = q[1] * q[2];
    const float wx = q[0] * q[3], wy = q[1] * q[3], wz = q[2] * q[3];
    M44 m;
    // transpose of the matrix written at src/quaternion.h:19-35
    m.m[0][0] = 1.0f - 2.0f * (yy + zz); m.m[1][0] = 2.0f * (xy + wz); m.m[2][0] = 2.0f * (xz - wy); m.m[3][0] = 0.0f;
    m.m[0][1] = 2.0f * (xy - wz); m.m[1][1] = 1.0f - 2.0f * (xx + zz); m.m[2][1] = 2.0f * (yz + wx); m.m[3][1] = 0.0f;
    m.m[0][2] = 2.0f * (xz + wy); m.m[1][2] = 2.0f * (yz - wx); m.m[2][2] = 1.0f - 2.0f * (xx + yy); m.m[3][2] = 0.0f;
    m.m[0][3] = t[0]; m.m[1][3] = t[1]; m.m[2][3] = t[2]; m.m[3][3] = 1.0f;
    return m;
}

struct ADScene {
    float useLightCoordinateSampling;
    M44 sampleToCam;     // row-major (deserialised from column-major storage)
    M44 camToWorld;
    float screenPixelCount, camDist;
    V3 bsphereCenter; float bsphereRadius;
};
LMC_HD ADScene ad_scene_deserialize(const float *s) {
    ADScene a;
    a.useLightCoordinateSampling = s[0];
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) a.sampleToCam.m[r][c] = s[1 + c * 4 + r];
    a.camToWorld = ser_static_matrix(s + 17);
    a.screenPixelCount = s[32]; a.camDist = s[33];
    a.bsphereCenter = ld3(s + 34); a.bsphereRadius = s[37];
    return a;
}

template <class T> struct ADRay { TV3<T> org, dir; };
template <class T> struct ADIsect { TV3<T> position, geomNormal, shadingNormal; };
template <class T> struct ADPathState {
    ADIsect<T> isect;
    TV3<T> wi;
    T accMISWPrev, accMISWThis;
    TV3<T> throughput;
};

// ---- shapes ------------------------------------------------------------------------------
// Intersect (static mode): TriangleIntersect<FloatType>, src/trianglemesh.cpp:81-105.  `st` is
// not produced: BSDF parameters arrive pre-evaluated in the buffer (textures are constants).
template <class T> LMC_HD_NOINLINE const float *ad_intersect(const float *buffer, const ADRay<T> &ray, ADIsect<T> &isect) {
    const float *b = buffer + 2;   // type, isMoving
    const V3 p0 = ld3(b), e1 = ld3(b + 3), e2 = ld3(b + 6), n0 = ld3(b + 9), n1 = ld3(b + 12), n2 = ld3(b + 15);
    isect.geomNormal = tv3c<T>(normalize(cross(e1, e2)));
    const TV3<T> e1t = tv3c<T>(e1), e2t = tv3c<T>(e2);
    const TV3<T> s1 = tcross(ray.dir, e2t);
    const T divisor = tdot(s1, e1t);
    const T invDivisor = ad_inverse(divisor);
    const TV3<T> s = ray.org - tv3c<T>(p0);
    const T u = tdot(s, s1) * invDivisor;
    const TV3<T> s2 = tcross(s, e1t);
    const T v = tdot(ray.dir, s2) * invDivisor;
    const T t = tdot(e2t, s2) * invDivisor;
    const T w = 1.0f - u - v;
    isect.position = ray.org + t * ray.dir;
    isect.shadingNormal = tnormalize(tscale(n0, w) + tscale(n1, u) + tscale(n2, v));
    return buffer + LMC_SER_SHAPE;
}
// SampleShape (static): SampleDirect<FloatType>, src/trianglemesh.cpp:313-327; pdf = invTotalArea
template <class T> LMC_HD_NOINLINE void ad_sample_shape(const float *buffer, const T &r0, const T &r1, TV3<T> &pos, TV3<T> &normal, float &pdf) {
    const float *b = buffer + 2;
    const V3 p0 = ld3(b), e1 = ld3(b + 3), e2 = ld3(b + 6), n0 = ld3(b + 9), n1 = ld3(b + 12), n2 = ld3(b + 15);
    const float adEps = 1e-6f;   // ADEpsilon<ADFloat>()
    const T a = ad_sqrt((1.0f + adEps) - r0);
    const T b1 = 1.0f - a;
    const T b2 = a * r1;
    pos = tv3c<T>(p0) + tscale(e1, b1) + tscale(e2, b2);
    normal = tnormalize(tscale(n0, 1.0f - b1 - b2) + tscale(n1, b1) + tscale(n2, b2));
    pdf = buffer[LMC_SER_SHAPE - 1];
}

// ---- BSDF twins ----------------------------------------------------------------------------
template <class T> LMC_HD_NOINLINE T ad_beckmann_D(const TV3<T> &localH, const T &alphaU, const T &alphaV) {
    const T cosTheta2 = ad_square(localH.z);
    const T e = (ad_square(localH.x) / ad_square(alphaU) + ad_square(localH.y) / ad_square(alphaV)) / cosTheta2;
    return ad_exp(-e) / (LMC_PI * alphaU * alphaV * ad_square(cosTheta2));
}
template <class T> LMC_HD_NOINLINE T ad_beckmann_G1(float alpha, const T &cosTheta) {
    const T tanTheta = ad_sqrt(ad_fabs((1.0f + 1e-6f) - ad_square(cosTheta))) / cosTheta;
    if (ad_val(tanTheta) <= 0.0f) return ad_const<T>(1.0f);
    const T a = ad_inverse(alpha * tanTheta);
    if (ad_val(a) >= 1.6f) return ad_const<T>(1.0f);
    const T aSqr = ad_square(a);
    return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
}
template <class T> LMC_HD_NOINLINE T ad_fresnel(const T &cosThetaI_, T &cosThetaT_, float eta, float invEta) {
    const float scale = (ad_val(cosThetaI_) > 0.0f) ? invEta : eta;
    const T cosThetaTSqr = 1.0f - (1.0f - ad_square(cosThetaI_)) * (scale * scale);
    if (ad_val(cosThetaTSqr) <= 0.0f) { cosThetaT_ = ad_const<T>(0.0f); return ad_const<T>(1.0f); }
    const T cosThetaI = ad_fabs(cosThetaI_);
    const T cosThetaT = ad_sqrt(cosThetaTSqr);
    const T etaCosThetaT = eta * cosThetaT;
    const T etaCosThetaI = eta * cosThetaI;
    const T Rs = (cosThetaI - etaCosThetaT) / (cosThetaI + etaCosThetaT);
    const T Rp = (etaCosThetaI - cosThetaT) / (etaCosThetaI + cosThetaT);
    cosThetaT_ = (ad_val(cosThetaI_) > 0.0f) ? -cosThetaT : cosThetaT;
    return 0.5f * (ad_square(Rs) + ad_square(Rp));
}
template <class T> LMC_HD TV3<T> ad_sample_cos_hemisphere(const T &r0, const T &r1) {
    const T phi = LMC_TWOPI * r0;
    const T tmp = ad_sqrt(ad_fmax(1.0f - r1, 1e-6f));
    T sp, cp; ad_sincos(phi, sp, cp);
    return tv3<T>(cp * tmp, sp * tmp, ad_sqrt(ad_fmax(r1, 1e-6f)));
}
template <class T> LMC_HD TV3<T> ad_sample_sphere(const T &c0, const T &c1, T &jacobian) {
    const T scaledTheta = LMC_TWOPI * c0;
    const T scaledPhi = LMC_PI * c1;
    T sinPhi, cosPhi, st, ct;
    ad_sincos(scaledPhi, sinPhi, cosPhi);
    ad_sincos(scaledTheta, st, ct);
    jacobian = ad_fabs(sinPhi) * LMC_TWOPI * LMC_PI;
    return tv3<T>(sinPhi * ct, sinPhi * st, cosPhi);
}

// normal flip shared by Lambertian / Phong twins (always two-sided in the AD code)
template <class T> LMC_HD void ad_face_normal(const TV3<T> &normal, T &cosWi, TV3<T> &n_) {
    if (ad_val(cosWi) > 0.0f) { n_ = normal; } else { n_ = -normal; cosWi = -cosWi; }
}

// buffer points at the BSDF record (type first)
template <class T> LMC_HD_NOINLINE void ad_evaluate_bsdf(bool adjoint, const float *buffer, const TV3<T> &wi, const TV3<T> &normal,
                                                const TV3<T> &wo, TV3<T> &contrib, T &cosWo, T &pdf, T &revPdf) {
    const int type = (int)buffer[0];
    const float *b = buffer + 1;
    if (type == BSDF_PHONG) {
        const V3 Kd = ld3(b), Ks = ld3(b + 3); const float exponent = b[6], KsWeight = b[7];
        T cosWi = tdot(normal, wi);
        TV3<T> n_; ad_face_normal(normal, cosWi, n_);
        cosWo = tdot(n_, wo);
        contrib = tv3c<T>(mk3s(0.0f)); pdf = ad_const<T>(0.0f); revPdf = ad_const<T>(0.0f);
        if (KsWeight > 0.0f) {
            const T alpha = tdot(treflect(wi, n_), wo);
            const T weight = ad_pow(alpha, exponent) * LMC_INVTWOPI;
            if (ad_val(weight) > 1e-10f) {
                contrib = tscale(Ks, (exponent + 2.0f) * weight);
                pdf = KsWeight * (exponent + 1.0f) * weight;
                revPdf = pdf;
            }
        }
        if (KsWeight < 1.0f) {
            const float tmp = (1.0f - KsWeight) * LMC_INVPI;
            contrib = contrib + tv3c<T>(Kd * LMC_INVPI);
            pdf = pdf + tmp * cosWo;
            revPdf = revPdf + tmp * cosWi;
        }
        contrib = contrib * cosWo;
    } else if (type == BSDF_ROUGHDIELECTRIC) {
        const V3 Ks = ld3(b), Kt = ld3(b + 3); const float eta = b[6], invEta = b[7], alpha = b[8];
        const T cosWi = tdot(wi, normal);
        cosWo = tdot(wo, normal);
        const bool reflect = ad_val(cosWi) * ad_val(cosWo) > 0.0f;
        const float eta_ = (ad_val(cosWi) > 0.0f) ? eta : invEta;
        const float revEta_ = (ad_val(cosWo) > 0.0f) ? eta : invEta;
        TV3<T> H = reflect ? tnormalize(wi + wo) : tnormalize(wi + tscalef(wo, eta_));
        if (ad_val(tdot(H, normal)) < 0.0f) H = -H;
        const T cosHWi = tdot(wi, H), cosHWo = tdot(wo, H);
        TV3<T> b0, b1; tcoordinate_system(normal, b0, b1);
        const TV3<T> localH = tv3<T>(tdot(b0, H), tdot(b1, H), tdot(normal, H));
        const T alphaT = ad_const<T>(alpha);
        const T D = ad_beckmann_D(localH, alphaT, alphaT);
        const T revCosHWi = cosHWo, revCosHWo = cosHWi;
        T dummy;
        const T F = ad_fresnel(cosHWi, dummy, eta, invEta);
        const T aCosWi = ad_fabs(cosWi), aCosWo = ad_fabs(cosWo);
        const T G = ad_beckmann_G1(alpha, aCosWi) * ad_beckmann_G1(alpha, aCosWo);
        const T scaledAlpha = alpha * (1.2f - 0.2f * ad_sqrt(aCosWi));
        const T prob = localH.z * ad_beckmann_D(localH, scaledAlpha, scaledAlpha);
        const T revScaledAlpha = alpha * (1.2f - 0.2f * ad_sqrt(aCosWo));
        const T revProb = localH.z * ad_beckmann_D(localH, revScaledAlpha, revScaledAlpha);
        if (reflect) {
            const T scalar = ad_fabs(F * D * G / (4.0f * cosWi));
            contrib = tscale(Ks, scalar);
            pdf = ad_fabs(prob * F / (4.0f * cosHWo));
            revPdf = ad_fabs(revProb * F / (4.0f * revCosHWo));
        } else {
            const T sqrtDenom = cosHWi + eta_ * cosHWo;
            const T revSqrtDenom = revCosHWi + revEta_ * revCosHWo;
            const float factor = adjoint ? 1.0f : square(inverse(eta_));
            const T scalar = ad_fabs(factor * ((1.0f - F) * D * G * square(eta_) * cosHWi * cosHWo) / (cosWi * ad_square(sqrtDenom)));
            contrib = tscale(Kt, scalar);
            pdf = ad_fabs(prob * (1.0f - F) * (square(eta_) * cosHWo) / ad_square(sqrtDenom));
            revPdf = ad_fabs(revProb * (1.0f - F) * (square(revEta_) * revCosHWo) / ad_square(revSqrtDenom));
        }
    } else if (type == BSDF_LAMBERTIAN) {
        const V3 Kd = ld3(b);
        T cosWi = tdot(normal, wi);
        TV3<T> n_; ad_face_normal(normal, cosWi, n_);
        cosWo = tdot(n_, wo);
        const T fwdScalar = cosWo * LMC_INVPI;
        contrib = tscale(Kd, fwdScalar);
        pdf = fwdScalar;
        revPdf = cosWi * LMC_INVPI;
    } else {
        contrib = tv3c<T>(mk3s(0.0f)); cosWo = ad_const<T>(0.0f); pdf = ad_const<T>(0.0f); revPdf = ad_const<T>(0.0f);
    }
}

template <class T> LMC_HD_NOINLINE void ad_sample_bsdf(bool adjoint, const float *buffer, const TV3<T> &wi, const TV3<T> &normal,
                                              const T &r0, const T &r1, float uDiscrete, TV3<T> &wo, TV3<T> &contrib,
                                              T &cosWo, T &pdf, T &revPdf) {
    const int type = (int)buffer[0];
    const float *b = buffer + 1;
    if (type == BSDF_ROUGHDIELECTRIC) {
        const V3 Ks = ld3(b), Kt = ld3(b + 3); const float eta = b[6], invEta = b[7], alpha = b[8];
        const T cosWi = tdot(wi, normal);
        const T scaledAlpha = alpha * (1.2f - 0.2f * ad_sqrt(ad_fabs(cosWi)));
        // SampleMicronormal<ADFloat>, src/microfacet.h:162-185
        const T phiM = LMC_TWOPI * r1;
        T sinPhiM, cosPhiM; ad_sincos(phiM, sinPhiM, cosPhiM);
        const T alphaSqr = ad_square(scaledAlpha);
        const T tanThetaMSqr = alphaSqr * (-ad_log(ad_fmax(1.0f - r0, 1e-6f)));
        const T cosThetaM = 1.0f / ad_sqrt(1.0f + tanThetaMSqr);
        const T cosThetaMSqr = ad_square(cosThetaM);
        const T mPdf = (1.0f - r0) / (LMC_PI * alphaSqr * cosThetaM * cosThetaMSqr);
        const T sinThetaM = ad_sqrt(ad_fmax(1.0f - cosThetaMSqr, 1e-6f));
        const TV3<T> localH = tv3<T>(sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM);
        TV3<T> b0, b1; tcoordinate_system(normal, b0, b1);
        const TV3<T> H = localH.x * b0 + localH.y * b1 + localH.z * normal;
        const T cosHWi = tdot(wi, H);
        T cosThetaT;
        const T F = ad_fresnel(cosHWi, cosThetaT, eta, invEta);
        TV3<T> refl;
        T cosHWo;
        if (uDiscrete <= ad_val(F)) {
            wo = treflect(wi, H);
            refl = tv3c<T>(Ks);
            cosHWo = tdot(wo, H);
            pdf = ad_fabs(mPdf * F / (4.0f * cosHWo));
            const T rev_dwh_dwo = ad_inverse(4.0f * cosHWi);
            cosWo = tdot(wo, normal);
            const T revScaledAlp = alpha * (1.2f - 0.2f * ad_sqrt(ad_fabs(cosWo)));
            const T revD = ad_beckmann_D(localH, revScaledAlp, revScaledAlp);
            revPdf = ad_fabs(F * revD * localH.z * rev_dwh_dwo);
        } else {
            wo = trefract(wi, H, cosThetaT, eta, invEta);
            const float eta_ = (ad_val(cosWi) > 0.0f) ? eta : invEta;
            const float factor = adjoint ? 1.0f : square(inverse(eta_));
            refl = tv3c<T>(Kt * factor);
            cosHWo = tdot(wo, H);
            const T sqrtDenom = cosHWi + eta_ * cosHWo;
            const T dwh_dwo = (square(eta_) * cosHWo) / ad_square(sqrtDenom);
            pdf = ad_fabs(mPdf * (1.0f - F) * ad_fabs(dwh_dwo));
            cosWo = tdot(wo, normal);
            const float revEta_ = (ad_val(cosWo) > 0.0f) ? eta : invEta;
            const T revSqrtDenom = cosHWo + revEta_ * cosHWi;
            const T rev_dwh_dwo = (square(revEta_) * cosHWi) / ad_square(revSqrtDenom);
            const T revScaledAlp = alpha * (1.2f - 0.2f * ad_sqrt(ad_fabs(cosWo)));
            const T revD = ad_beckmann_D(localH, revScaledAlp, revScaledAlp);
            revPdf = ad_fabs((1.0f - F) * revD * localH.z * rev_dwh_dwo);
        }
        const T aCosWi = ad_fabs(cosWi), aCosWo = ad_fabs(cosWo);
        const T alphaT = ad_const<T>(alpha);
        const T D = ad_beckmann_D(localH, alphaT, alphaT);
        const T G = ad_beckmann_G1(alpha, aCosWi) * ad_beckmann_G1(alpha, aCosWo);
        const T numerator = D * G * cosHWi;
        const T denominator = mPdf * aCosWi;
        contrib = refl * ad_fabs(numerator / denominator);
    } else if (type == BSDF_PHONG) {
        const V3 Kd = ld3(b), Ks = ld3(b + 3); const float exponent = b[6], KsWeight = b[7];
        T cosWi = tdot(normal, wi);
        TV3<T> n_; ad_face_normal(normal, cosWi, n_);
        const TV3<T> R = treflect(wi, n_);
        if (uDiscrete > KsWeight) {
            const TV3<T> localDir = ad_sample_cos_hemisphere(r0, r1);
            TV3<T> b0, b1; tcoordinate_system(n_, b0, b1);
            wo = localDir.x * b0 + localDir.y * b1 + localDir.z * n_;
        } else {
            const float power = 1.0f / (exponent + 1.0f);
            const T cosAlpha = ad_pow(r1, power);
            const T sinAlpha = ad_sqrt(ad_fmax(1.0f - ad_square(cosAlpha), 1e-6f));
            const T phi = LMC_TWOPI * r0;
            T sphi, cphi; ad_sincos(phi, sphi, cphi);
            const TV3<T> localDir = tv3<T>(sinAlpha * cphi, sinAlpha * sphi, cosAlpha);
            TV3<T> b0, b1; tcoordinate_system(R, b0, b1);
            wo = localDir.x * b0 + localDir.y * b1 + localDir.z * R;
        }
        cosWo = tdot(n_, wo);
        contrib = tv3c<T>(mk3s(0.0f)); pdf = ad_const<T>(0.0f); revPdf = ad_const<T>(0.0f);
        if (KsWeight > 0.0f) {
            const T alpha = tdot(R, wo);
            const T weight = ad_pow(alpha, exponent) * LMC_INVTWOPI;
            if (ad_val(weight) > 1e-10f) {
                contrib = tscale(Ks, (exponent + 2.0f) * weight);
                pdf = KsWeight * (exponent + 1.0f) * weight;
                revPdf = pdf;
            }
        }
        if (KsWeight < 1.0f) {
            const float tmp = (1.0f - KsWeight) * LMC_INVPI;
            contrib = contrib + tv3c<T>(Kd * LMC_INVPI);
            pdf = pdf + tmp * cosWo;
            revPdf = revPdf + tmp * cosWi;
        }
        contrib = contrib * cosWo;
        contrib = contrib * ad_inverse(pdf);
    } else if (type == BSDF_LAMBERTIAN) {
        const V3 Kd = ld3(b);
        T cosWi = tdot(wi, normal);
        TV3<T> n_; ad_face_normal(normal, cosWi, n_);
        TV3<T> b0, b1; tcoordinate_system(n_, b0, b1);
        const TV3<T> r = ad_sample_cos_hemisphere(r0, r1);
        wo = r.x * b0 + r.y * b1 + r.z * n_;
        cosWo = r.z;
        pdf = r.z * LMC_INVPI;
        contrib = tv3c<T>(Kd);
        revPdf = cosWi * LMC_INVPI;
    } else {
        wo = tv3c<T>(mk3s(0.0f)); contrib = tv3c<T>(mk3s(0.0f));
        cosWo = ad_const<T>(0.0f); pdf = ad_const<T>(0.0f); revPdf = ad_const<T>(0.0f);
    }
}

template <class T> LMC_HD T ad_shading_normal_correction_adjoint(const TV3<T> &wi, const ADIsect<T> &isect, const TV3<T> &wo) {
    const T cosWi = tdot(isect.shadingNormal, wi);
    const T cosWo = tdot(isect.shadingNormal, wo);
    const T wiDotGeoN = tdot(isect.geomNormal, wi);
    const T woDotGeoN = tdot(isect.geomNormal, wo);
    return ad_fabs((woDotGeoN * cosWi) / (wiDotGeoN * cosWo));
}

// ---- light twins -----------------------------------------------------------------------------
template <class T> LMC_HD T ad_tent(const T &s) {
    if (ad_val(s) < 0.5f) return 1.0f - ad_sqrt(2.0f * s);
    return ad_sqrt(2.0f * (s - 0.5f)) - 1.0f;
}
// env light record: buffer points at the type; fields follow contiguously (the transforms take
// 15 floats each; the 2 spare floats of the 56-float record are at its end)
struct ADEnvRec {
    M44 toWorld, toLight;
    float cdfCol0, cdfCol1, cdfRow0, cdfRow1, col, row, pixelSize[2];
    V3 img00, img10, img01, img11;
    float rowWeight0, rowWeight1, normalization;
};
LMC_HD_NOINLINE ADEnvRec ad_env_deserialize(const float *buffer) {
    ADEnvRec e;
    const float *b = buffer + 1;
    e.toWorld = ser_static_matrix(b); e.toLight = ser_static_matrix(b + 15);
    b += 30;
    e.cdfCol0 = b[0]; e.cdfCol1 = b[1]; e.cdfRow0 = b[2]; e.cdfRow1 = b[3]; e.col = b[4]; e.row = b[5];
    e.pixelSize[0] = b[6]; e.pixelSize[1] = b[7];
    e.img00 = ld3(b + 8); e.img10 = ld3(b + 11); e.img01 = ld3(b + 14); e.img11 = ld3(b + 17);
    e.rowWeight0 = b[20]; e.rowWeight1 = b[21]; e.normalization = b[22];
    return e;
}
template <class T> LMC_HD_NOINLINE void ad_env_sample_direction(const ADEnvRec &e, const T &r0, const T &r1, TV3<T> &dirToLight,
                                                       TV3<T> &value, T &pdf) {
    const T u0 = (r0 - e.cdfCol0) / (e.cdfCol1 - e.cdfCol0);
    const T u1 = (r1 - e.cdfRow0) / (e.cdfRow1 - e.cdfRow0);
    const T tx = ad_tent(u0), ty = ad_tent(u1);
    const T plx = e.col + tx, ply = e.row + ty;
    const T phi = (plx + 0.5f) * e.pixelSize[0];
    const T theta = (ply + 0.5f) * e.pixelSize[1];
    T sinPhi, cosPhi, sinTheta, cosTheta;
    ad_sincos(phi, sinPhi, cosPhi);
    ad_sincos(theta, sinTheta, cosTheta);
    dirToLight = txform_vector(e.toWorld, tv3<T>(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta));
    const T dx1 = tx, dx2 = 1.0f - tx, dy1 = ty, dy2 = 1.0f - ty;
    const TV3<T> value1 = tscale(e.img00, dx2) * dy2 + tscale(e.img10, dx1) * dy2;
    const TV3<T> value2 = tscale(e.img01, dx2) * dy1 + tscale(e.img11, dx1) * dy1;
    value = value1 + value2;
    pdf = (tluminance(value1) * e.rowWeight0 + tluminance(value2) * e.rowWeight1) * e.normalization /
          ad_fmax(ad_fabs(sinTheta), 1e-7f);
}

// SampleDirect twin; buffer points at the light record
template <class T> LMC_HD_NOINLINE void ad_sample_direct(const float *buffer, const ADScene &scn, const TV3<T> &pos, const T &r0,
                                                const T &r1, TV3<T> &dirToLight, TV3<T> &lightContrib, T &cosAtLight,
                                                T &directPdf, T &emissionPdf) {
    const int type = (int)buffer[0];
    if (type == LIGHT_POINT) {
        const V3 lightPos = ld3(buffer + 1), emission = ld3(buffer + 4);
        dirToLight = tv3c<T>(lightPos) - pos;
        const T distSq = tlength_squared(dirToLight);
        directPdf = distSq;
        const T dist = ad_sqrt(distSq);
        dirToLight = dirToLight * ad_inverse(dist);   // dirToLight / dist
        lightContrib = tscale(emission, ad_inverse(distSq));
        emissionPdf = ad_const<T>(1.0f / (4.0f * LMC_PI));
        cosAtLight = ad_const<T>(1.0f);
    } else if (type == LIGHT_AREA) {
        TV3<T> posOnLight, normalOnLight; float shapePdf;
        ad_sample_shape(buffer + 1, r0, r1, posOnLight, normalOnLight, shapePdf);
        const V3 emission = ld3(buffer + 1 + LMC_SER_SHAPE);
        dirToLight = posOnLight - pos;
        const T distSq = tlength_squared(dirToLight);
        const T dist = ad_sqrt(distSq);
        dirToLight = dirToLight * ad_inverse(dist);
        cosAtLight = -tdot(dirToLight, normalOnLight);
        directPdf = shapePdf * distSq / cosAtLight;
        lightContrib = tscale(emission, ad_inverse(directPdf));
        emissionPdf = shapePdf * cosAtLight * LMC_INVPI;
    } else {
        const ADEnvRec e = ad_env_deserialize(buffer);
        TV3<T> value;
        ad_env_sample_direction(e, r0, r1, dirToLight, value, directPdf);
        lightContrib = value * ad_inverse(directPdf);
        cosAtLight = ad_const<T>(1.0f);
        const float positionPdf = LMC_INVPI / square(scn.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
    }
}
// Emission twin
template <class T> LMC_HD_NOINLINE void ad_emission(const float *buffer, const ADScene &scn, const TV3<T> &dirToLight,
                                           const TV3<T> &normalOnLight, TV3<T> &emission, T &directPdf, T &emissionPdf) {
    const int type = (int)buffer[0];
    if (type == LIGHT_AREA) {
        const float shapePdf = buffer[1 + LMC_SER_SHAPE - 1];
        const V3 em = ld3(buffer + 1 + LMC_SER_SHAPE);
        const T cosAtLight = -tdot(normalOnLight, dirToLight);
        emission = tv3c<T>(em);
        directPdf = ad_const<T>(shapePdf);
        emissionPdf = cosAtLight * directPdf * LMC_INVPI;
    } else if (type == LIGHT_ENV) {
        const ADEnvRec e = ad_env_deserialize(buffer);
        const TV3<T> d = txform_vector(e.toLight, dirToLight);
        const T uvx = ad_atan2(d.x, -d.z) / e.pixelSize[0] - 0.5f;
        const T uvy = ad_acos(d.y) / e.pixelSize[1] - 0.5f;
        const T dx1 = uvx - e.col, dx2 = 1.0f - dx1, dy1 = uvy - e.row, dy2 = 1.0f - dy1;
        const TV3<T> value1 = tscale(e.img00, dx2) * dy2 + tscale(e.img10, dx1) * dy2;
        const TV3<T> value2 = tscale(e.img01, dx2) * dy1 + tscale(e.img11, dx1) * dy1;
        emission = value1 + value2;
        const T sinTheta = ad_sqrt(ad_fmax(1.0f - ad_square(d.y), 1e-6f));
        directPdf = (tluminance(value1) * e.rowWeight0 + tluminance(value2) * e.rowWeight1) * e.normalization /
                    ad_fmax(ad_fabs(sinTheta), 1e-7f);
        const float positionPdf = LMC_INVPI / square(scn.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
    } else {
        emission = tv3c<T>(mk3s(0.0f)); directPdf = ad_const<T>(0.0f); emissionPdf = ad_const<T>(0.0f);
    }
}
template <class T> LMC_HD void ad_sample_concentric_disc(const T &r0, const T &r1, T &ox, T &oy) {
    const T a1 = 2.0f * r0 - 1.0f, a2 = 2.0f * r1 - 1.0f;
    T r, phi;
    if (ad_val(a1) == 0.0f || ad_val(a2) == 0.0f) { r = ad_const<T>(0.0f); phi = ad_const<T>(0.0f); }
    else if (ad_val(a1) * ad_val(a1) > ad_val(a2) * ad_val(a2)) { r = a1; phi = LMC_PIOVERFOUR * (a2 / a1); }
    else { r = a2; phi = LMC_PIOVERTWO - (a1 / a2) * LMC_PIOVERFOUR; }
    T sphi, cphi; ad_sincos(phi, sphi, cphi);
    ox = r * cphi; oy = r * sphi;
}
// Emit twin
template <class T> LMC_HD_NOINLINE void ad_emit(const float *buffer, const ADScene &scn, const T &p0, const T &p1, const T &d0,
                                       const T &d1, ADRay<T> &ray, TV3<T> &emission, T &cosAtLight, T &emissionPdf,
                                       T &directPdf) {
    const int type = (int)buffer[0];
    if (type == LIGHT_POINT) {
        const V3 lightPos = ld3(buffer + 1), em = ld3(buffer + 4);
        ray.org = tv3c<T>(lightPos);
        T jac; ray.dir = ad_sample_sphere(d0, d1, jac);
        emission = tv3c<T>(em);
        emissionPdf = ad_const<T>(1.0f / (4.0f * LMC_PI));
        cosAtLight = ad_const<T>(1.0f); directPdf = ad_const<T>(1.0f);
    } else if (type == LIGHT_AREA) {
        TV3<T> normalOnLight; float shapePdf;
        ad_sample_shape(buffer + 1, p0, p1, ray.org, normalOnLight, shapePdf);
        const V3 em = ld3(buffer + 1 + LMC_SER_SHAPE);
        const TV3<T> d = ad_sample_cos_hemisphere(d0, d1);
        TV3<T> b0, b1; tcoordinate_system(normalOnLight, b0, b1);
        ray.dir = d.x * b0 + d.y * b1 + d.z * normalOnLight;
        emission = tv3c<T>(em * ((float)M_PI / shapePdf));
        cosAtLight = d.z;
        emissionPdf = d.z * LMC_INVPI * shapePdf;
        directPdf = ad_const<T>(shapePdf);
    } else {
        const ADEnvRec e = ad_env_deserialize(buffer);
        ad_env_sample_direction(e, d0, d1, ray.dir, emission, directPdf);
        ray.dir = -ray.dir;
        T ox, oy; ad_sample_concentric_disc(p0, p1, ox, oy);
        TV3<T> b0, b1; tcoordinate_system(ray.dir, b0, b1);
        const TV3<T> perpOffset = ox * b0 + oy * b1;
        ray.org = tv3c<T>(scn.bsphereCenter) + tscalef(perpOffset - ray.dir, scn.bsphereRadius);
        cosAtLight = ad_const<T>(1.0f);
        const float positionPdf = LMC_INVPI / square(scn.bsphereRadius);
        emissionPdf = directPdf * positionPdf;
    }
}

// ---- camera twin (static) -------------------------------------------------------------------
template <class T> LMC_HD void ad_sample_primary(const ADScene &scn, const T &sx, const T &sy, ADRay<T> &ray) {
    const TV3<T> o = txform_point(scn.sampleToCam, tv3<T>(sx, sy, ad_const<T>(0.0f)));
    const TV3<T> dir = tnormalize(o);
    ray.org = tv3c<T>(xform_point(scn.camToWorld, mk3s(0.0f)));
    ray.dir = txform_vector(scn.camToWorld, dir);
}

template <class T> LMC_HD void ad_convert_mis(const ADRay<T> &ray, ADPathState<T> &ps) {
    ps.accMISWPrev = ps.accMISWPrev * ad_mis(tdistance_squared(ray.org, ps.isect.position));
    const T invCosTheta = ad_inverse(ad_mis(ad_fabs(tdot(ray.dir, ps.isect.shadingNormal))));
    ps.accMISWPrev = ps.accMISWPrev * invCosTheta;
    ps.accMISWThis = ps.accMISWThis * invCosTheta;
}

// BSDFSampling<adjoint, fixedDiscrete = false> without light-coordinate sampling
template <class T> LMC_HD_NOINLINE const float *ad_bsdf_sampling(bool adjoint, const float *buffer, const T &r0, const T &r1,
                                                        float bsdfDiscrete, float useAbsoluteParam, ADPathState<T> &ps,
                                                        TV3<T> &dir) {
    TV3<T> bsdfContrib; T cosWo, bsdfPdf, bsdfRevPdf, jacobian;
    if (useAbsoluteParam == 0.0f) {
        ad_sample_bsdf(adjoint, buffer, ps.wi, ps.isect.shadingNormal, r0, r1, bsdfDiscrete, dir, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf);
        jacobian = ad_const<T>(1.0f);
    } else {
        dir = ad_sample_sphere(r0, r1, jacobian);
        ad_evaluate_bsdf(adjoint, buffer, ps.wi, ps.isect.shadingNormal, dir, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf);
    }
    if (adjoint) {
        const T factor = ad_shading_normal_correction_adjoint(ps.wi, ps.isect, dir);
        bsdfContrib = bsdfContrib * factor;
    }
    bsdfContrib = bsdfContrib * jacobian;
    ps.accMISWThis = ad_mis(cosWo / bsdfPdf) * (ps.accMISWThis * ad_mis(bsdfRevPdf) + ps.accMISWPrev);
    ps.accMISWPrev = ad_mis(ad_inverse(bsdfPdf));
    ps.throughput = tcmul(ps.throughput, bsdfContrib);
    return buffer + LMC_SER_BSDF;
}

// The path function: log(Luminance(contrib)) of a (camDepth, lightDepth) path.
// primary: D+1 values (time first); `pss` carries primary[1..D] as T (seeded duals or floats).
// Lockstep hook.  The evaluator is ~20 k straight-line instructions that a warp walks once per sweep, so
// the device kernels are bound by instruction fetch, not by arithmetic.  Its control flow between the
// marks below depends ONLY on (maxCamDepth, maxLightDepth): when every thread of a block evaluates the
// same path class (cuda/chain_kernels.cuh builds class-pure blocks; k_eval_batch gets the class as a
// kernel argument), a block-wide barrier at each mark keeps the block's warps inside one
// instruction-cache window and the fetched code is shared instead of streamed per warp (measured:
// gradient kernel 4.8 -> 2.9 ms per iteration of 2^20 chains).  Barriers change no arithmetic; the
// host twin compiles them away.
#if defined(__CUDA_ARCH__) && !defined(LMC_NO_GRAD_LOCKSTEP)
#define LMC_VERTEX_SYNC() __syncthreads()
#else
#define LMC_VERTEX_SYNC() ((void)0)
#endif
#define LMC_VERTEX_SYNC2() LMC_VERTEX_SYNC()

// PSS: anything indexable that yields the i-th primary-sample value as a T: a float pointer for the forward
// value, a seed object for the derivative sweeps (the seeded duals are built where they are consumed instead
// of being staged in a local array: 100 local-memory stores fewer per sweep)
template <class T, class PSS>
LMC_HD_NOINLINE T eval_path_loglum(int maxCamDepth, int maxLightDepth, const float *sceneBuf, const float *vertParams, const PSS &pss) {
    const ADScene scn = ad_scene_deserialize(sceneBuf);
    const float *buffer = vertParams + 3;   // lensVertexPos
    int pi = 0;
    const float *lgtBSDFBuffer = nullptr;
    ADPathState<T> lps;
    TV3<T> contrib = tv3c<T>(mk3s(0.0f));
    if (maxLightDepth > 1) {
        const float lightPickProb = *buffer++;
        ADRay<T> ray;
        const T p0 = pss[pi], p1 = pss[pi + 1], d0 = pss[pi + 2], d1 = pss[pi + 3];
        pi += 4;
        const int lightType = (int)buffer[0];
        {   // EmitFromLight
            T cosLight, emissionPdf, directPdf;
            ad_emit(buffer, scn, p0, p1, d0, d1, ray, lps.throughput, cosLight, emissionPdf, directPdf);
            buffer += LMC_SER_LIGHT;
            emissionPdf = emissionPdf * lightPickProb;
            directPdf = directPdf * lightPickProb;
            lps.throughput = tscalef(lps.throughput, inverse(lightPickProb));
            lps.accMISWPrev = ad_mis(directPdf / emissionPdf);
            // sic: the AD twin tests `== PointLight`, the tracer `!IsDelta()`
            if (lightType == LIGHT_POINT) lps.accMISWThis = ad_mis(cosLight / emissionPdf);
            else lps.accMISWThis = ad_const<T>(0.0f);
        }
        for (int lgtDepth = 0; lgtDepth < maxLightDepth - 1; lgtDepth++) {
            LMC_VERTEX_SYNC();
            buffer = ad_intersect(buffer, ray, lps.isect);
            const float bsdfDiscrete = *buffer++;
            const float useAbsoluteParam = *buffer++;
            lps.wi = -ray.dir;
            if (lgtDepth == 0) {   // ConvertMISLightEmit
                const T invCosTheta = ad_inverse(ad_mis(ad_fabs(tdot(ray.dir, lps.isect.shadingNormal))));
                if (lightType == LIGHT_ENV) lps.accMISWPrev = lps.accMISWPrev * (invCosTheta * 1.0f);
                else lps.accMISWPrev = lps.accMISWPrev * (invCosTheta * ad_mis(tdistance_squared(ray.org, lps.isect.position)));
                lps.accMISWThis = lps.accMISWThis * invCosTheta;
            } else {
                ad_convert_mis(ray, lps);
            }
            if (lgtDepth == maxLightDepth - 2) {
                if (maxCamDepth == 1) {   // ConnectToCamera
                    ADRay<T> centerRay;
                    ad_sample_primary(scn, ad_const<T>(0.5f), ad_const<T>(0.5f), centerRay);
                    TV3<T> dirToCamera = centerRay.org - lps.isect.position;
                    const T distSq = tlength_squared(dirToCamera);
                    const T dist = ad_sqrt(distSq);
                    dirToCamera = dirToCamera * ad_inverse(dist);
                    TV3<T> bsdfContrib; T cosToCamera, bsdfPdf, bsdfRevPdf;
                    ad_evaluate_bsdf(true, buffer, lps.wi, lps.isect.shadingNormal, dirToCamera, bsdfContrib, cosToCamera, bsdfPdf, bsdfRevPdf);
                    const T factor = ad_shading_normal_correction_adjoint(lps.wi, lps.isect, dirToCamera);
                    bsdfContrib = bsdfContrib * factor;
                    const T invCosAtCamera = -ad_inverse(tdot(centerRay.dir, dirToCamera));
                    const T imagePointToCameraDist = scn.camDist * invCosAtCamera;
                    const T imageToSolidAngleFactor = ad_square(imagePointToCameraDist) * invCosAtCamera;
                    const T imageToSurfaceFactor = imageToSolidAngleFactor * ad_fabs(cosToCamera) / distSq;
                    const T wLight = ad_mis(imageToSurfaceFactor / scn.screenPixelCount) * (lps.accMISWPrev + lps.accMISWThis * ad_mis(bsdfRevPdf));
                    const T misWeight = ad_inverse(wLight + 1.0f);
                    const T surfaceToImageFactor = cosToCamera / imageToSurfaceFactor;
                    const TV3<T> c = (misWeight * bsdfContrib) * ad_inverse(scn.screenPixelCount * surfaceToImageFactor);
                    lps.throughput = tcmul(c, lps.throughput);
                    contrib = lps.throughput;
                }
                lgtBSDFBuffer = buffer;
                buffer += LMC_SER_BSDF;
                break;
            }
            const T r0 = pss[pi], r1 = pss[pi + 1];
            pi += 2;
            LMC_VERTEX_SYNC2();
            buffer = ad_bsdf_sampling(true, buffer, r0, r1, bsdfDiscrete, useAbsoluteParam, lps, ray.dir);
            const float rrWeight = *buffer++;
            lps.throughput = tscalef(lps.throughput, rrWeight);
            ray.org = lps.isect.position;
        }
    }
    if (maxCamDepth > 1) {
        const T sx = pss[pi], sy = pss[pi + 1];
        pi += 2;
        ADRay<T> ray;
        ADPathState<T> cps;
        {   // EmitFromCamera
            ADRay<T> centerRay;
            ad_sample_primary(scn, ad_const<T>(0.5f), ad_const<T>(0.5f), centerRay);
            ad_sample_primary(scn, sx, sy, ray);
            const T cosAtCamera = tdot(centerRay.dir, ray.dir);
            const T imagePointToCameraDist = scn.camDist / cosAtCamera;
            const T cameraPdf = ad_square(imagePointToCameraDist) / cosAtCamera;
            cps.throughput = tv3c<T>(mk3s(1.0f));
            cps.accMISWPrev = ad_mis(scn.screenPixelCount / cameraPdf);
            cps.accMISWThis = ad_const<T>(0.0f);
        }
        for (int camDepth = 0; camDepth < maxCamDepth - 1; camDepth++) {
            LMC_VERTEX_SYNC();
            buffer = ad_intersect(buffer, ray, cps.isect);
            cps.wi = -ray.dir;
            if (camDepth == maxCamDepth - 2 && maxLightDepth == 0) {
                const int lightType = (int)buffer[0];
                // ConvertMISLightHit
                if (lightType != LIGHT_ENV) {
                    const T distSq = ad_mis(tdistance_squared(ray.org, cps.isect.position));
                    const T invCosTheta = ad_inverse(ad_mis(ad_fabs(tdot(ray.dir, cps.isect.shadingNormal))));
                    cps.accMISWPrev = cps.accMISWPrev * (invCosTheta * distSq);
                    cps.accMISWThis = cps.accMISWThis * invCosTheta;
                }
                // HandleHitLight
                TV3<T> emission; T directPdf, emissionPdf;
                ad_emission(buffer, scn, ray.dir, cps.isect.shadingNormal, emission, directPdf, emissionPdf);
                buffer += LMC_SER_LIGHT;
                cps.throughput = tcmul(cps.throughput, emission);
                const float lightPickProb = *buffer++;
                directPdf = directPdf * lightPickProb;
                emissionPdf = emissionPdf * lightPickProb;
                const T wCamera = ad_mis(directPdf) * cps.accMISWPrev + ad_mis(emissionPdf) * cps.accMISWThis;
                const T misWeight = ad_inverse(1.0f + wCamera);
                cps.throughput = cps.throughput * misWeight;
                contrib = cps.throughput;
                break;
            }
            ad_convert_mis(ray, cps);
            if (camDepth == maxCamDepth - 2) {
                LMC_VERTEX_SYNC2();
                if (maxLightDepth == 1) {   // DirectLighting
                    const T r0 = pss[pi], r1 = pss[pi + 1];
                    pi += 2;
                    const int lightType = (int)buffer[0];
                    TV3<T> dirToLight, lightContrib; T cosAtLight, directPdf, emissionPdf;
                    ad_sample_direct(buffer, scn, cps.isect.position, r0, r1, dirToLight, lightContrib, cosAtLight, directPdf, emissionPdf);
                    buffer += LMC_SER_LIGHT;
                    TV3<T> bsdfContrib; T cosToLight, bsdfPdf, bsdfRevPdf;
                    ad_evaluate_bsdf(false, buffer, cps.wi, cps.isect.shadingNormal, dirToLight, bsdfContrib, cosToLight, bsdfPdf, bsdfRevPdf);
                    buffer += LMC_SER_BSDF;
                    const float lightPickProb = *buffer++;
                    cps.throughput = tcmul(cps.throughput, bsdfContrib);
                    cps.throughput = tscalef(tcmul(cps.throughput, lightContrib), inverse(lightPickProb));
                    T wLight;
                    if (lightType == LIGHT_POINT) wLight = ad_const<T>(0.0f);
                    else wLight = ad_mis(bsdfPdf / (lightPickProb * directPdf));
                    const T wCamera = ad_mis(emissionPdf * cosToLight / (directPdf * cosAtLight)) *
                                      (cps.accMISWPrev + cps.accMISWThis * ad_mis(bsdfRevPdf));
                    const T misWeight = ad_inverse(wLight + 1.0f + wCamera);
                    cps.throughput = cps.throughput * misWeight;
                } else {   // ConnectVertex
                    TV3<T> dirToLight = lps.isect.position - cps.isect.position;
                    const T distSq = tlength_squared(dirToLight);
                    const T dist = ad_sqrt(distSq);
                    dirToLight = dirToLight * ad_inverse(dist);
                    TV3<T> camBsdfFactor; T cosCamera, camBsdfPdf, camBsdfRevPdf;
                    ad_evaluate_bsdf(false, buffer, cps.wi, cps.isect.shadingNormal, dirToLight, camBsdfFactor, cosCamera, camBsdfPdf, camBsdfRevPdf);
                    TV3<T> lgtBsdfFactor; T cosLight, lgtBsdfPdf, lgtBsdfRevPdf;
                    const TV3<T> negDir = -dirToLight;
                    ad_evaluate_bsdf(true, lgtBSDFBuffer, lps.wi, lps.isect.shadingNormal, negDir, lgtBsdfFactor, cosLight, lgtBsdfPdf, lgtBsdfRevPdf);
                    const T lgtFactor = ad_shading_normal_correction_adjoint(lps.wi, lps.isect, negDir);
                    lgtBsdfFactor = lgtBsdfFactor * lgtFactor;
                    const T geometryTerm = ad_inverse(distSq);
                    const T camBsdfDirPdfA = camBsdfPdf * cosLight * geometryTerm;
                    const T lgtBsdfDirPdfA = lgtBsdfPdf * cosCamera * geometryTerm;
                    const T wLight = ad_mis(camBsdfDirPdfA) * (lps.accMISWPrev + lps.accMISWThis * ad_mis(lgtBsdfRevPdf));
                    const T wCamera = ad_mis(lgtBsdfDirPdfA) * (cps.accMISWPrev + cps.accMISWThis * ad_mis(camBsdfRevPdf));
                    const T misWeight = ad_inverse(wLight + 1.0f + wCamera);
                    cps.throughput = tcmul(lps.throughput, cps.throughput);
                    cps.throughput = tcmul(cps.throughput, camBsdfFactor);
                    cps.throughput = tcmul(cps.throughput, lgtBsdfFactor) * (geometryTerm * misWeight);
                }
                contrib = cps.throughput;
                break;
            }
            const T r0 = pss[pi], r1 = pss[pi + 1];
            pi += 2;
            const float bsdfDiscrete = *buffer++;
            const float useAbsoluteParam = *buffer++;
            LMC_VERTEX_SYNC2();
            buffer = ad_bsdf_sampling(false, buffer, r0, r1, bsdfDiscrete, useAbsoluteParam, cps, ray.dir);
            const float rrWeight = *buffer++;
            cps.throughput = tscalef(cps.throughput, rrWeight);
            ray.org = cps.isect.position;
        }
    }
    return ad_log(tluminance(contrib));
}

// Forward value only.
LMC_HD_NOINLINE float path_loglum(int camDepth, int lightDepth, const float *sceneBuf, const float *primary, const float *vertParams) {
    const float *pss = primary + 1;
    return eval_path_loglum<float>(camDepth, lightDepth, sceneBuf, vertParams, pss);
}

// Gradient w.r.t. primary[1..D] (time excluded: Static mode), NCHUNK directions per sweep.
#ifndef LMC_GRAD_CHUNK
#define LMC_GRAD_CHUNK 4
#endif
#define LMC_GRAD_MAXDIM 24
// seeds of one first-order sweep: value primary[1 + i], derivative e_(i - base) for the chunk's directions
struct GradSeed {
    const float *primary; int base;
    LMC_HD Dual<LMC_GRAD_CHUNK> operator[](int i) const {
        Dual<LMC_GRAD_CHUNK> r;
        r.v = primary[1 + i];
        for (int k = 0; k < LMC_GRAD_CHUNK; k++) r.d[k] = (i == base + k) ? 1.0f : 0.0f;
        return r;
    }
};
LMC_HD_NOINLINE float path_loglum_grad(int camDepth, int lightDepth, const float *sceneBuf, const float *primary,
                                       const float *vertParams, float *grad) {
    const int dim = 2 * ((camDepth + lightDepth - 1) > 2 ? (camDepth + lightDepth - 1) : 2);
    typedef Dual<LMC_GRAD_CHUNK> D;
    float value = 0.0f;
    for (int base = 0; base < dim; base += LMC_GRAD_CHUNK) {
        GradSeed pss; pss.primary = primary; pss.base = base;
        const D r = eval_path_loglum<D>(camDepth, lightDepth, sceneBuf, vertParams, pss);
        value = r.v;
        for (int k = 0; k < LMC_GRAD_CHUNK; k++) if (base + k < dim) grad[base + k] = r.d[k];
    }
    return value;
}

// Gradient AND Hessian w.r.t. primary[1..D] by second-order forward mode: nested duals carry
// LMC_HESS_CHUNK outer x LMC_HESS_CHUNK inner directions per sweep; the chunk pairs (a <= b) fill
// the symmetric matrix.  Replaces the forward-over-reverse code `evaluate_path_bidir_<c>_<l>_static_derv`
// (src/chad.cpp:333-545); hess is row-major D x D (hess[i * D + j] = d2 f / dx_i dx_j, symmetric).
#define LMC_HESS_CHUNK 2
// Which Hessian the H2MC mutation runs: > 0 = forward-over-reverse (pathgrad_rev.h path_loglum_hess_rev) with that many
// directions per reverse sweep, 0 = the second-order forward evaluator below.  Measured on B200 (torus, maxdepth 8,
// 2^20 chains, Hessian kernels per iteration): second-order forward 62 ms, forward-over-reverse 1 direction 30 ms,
// 2 directions 24 ms (ptxas: 75 s / 153 s for the kernel).
#ifndef LMC_HESS_REV_CHUNK
#define LMC_HESS_REV_CHUNK 2
#endif
#define LMC_HESS_MAXDIM 16
// seeds of one second-order sweep: outer directions ba.., inner directions bb..
struct HessSeed {
    const float *primary; int ba, bb;
    LMC_HD DualT<DualT<float, LMC_HESS_CHUNK>, LMC_HESS_CHUNK> operator[](int i) const {
        DualT<DualT<float, LMC_HESS_CHUNK>, LMC_HESS_CHUNK> r;
        r.v.v = primary[1 + i];
        for (int k = 0; k < LMC_HESS_CHUNK; k++) r.v.d[k] = (i == bb + k) ? 1.0f : 0.0f;
        for (int j = 0; j < LMC_HESS_CHUNK; j++) {
            r.d[j].v = (i == ba + j) ? 1.0f : 0.0f;
            for (int k = 0; k < LMC_HESS_CHUNK; k++) r.d[j].d[k] = 0.0f;
        }
        return r;
    }
};
LMC_HD_NOINLINE float path_loglum_hess(int camDepth, int lightDepth, const float *sceneBuf, const float *primary,
                                       const float *vertParams, float *grad, float *hess) {
    const int dim = 2 * ((camDepth + lightDepth - 1) > 2 ? (camDepth + lightDepth - 1) : 2);
    typedef DualT<float, LMC_HESS_CHUNK> D1;
    typedef DualT<D1, LMC_HESS_CHUNK> D2;
    float value = 0.0f;
    for (int ba = 0; ba < dim; ba += LMC_HESS_CHUNK) {
        for (int bb = ba; bb < dim; bb += LMC_HESS_CHUNK) {
            HessSeed pss; pss.primary = primary; pss.ba = ba; pss.bb = bb;
            const D2 r = eval_path_loglum<D2>(camDepth, lightDepth, sceneBuf, vertParams, pss);
            value = r.v.v;
            for (int j = 0; j < LMC_HESS_CHUNK; j++) {
                if (ba + j >= dim) continue;
                if (bb == ba) grad[ba + j] = r.d[j].v;
                for (int k = 0; k < LMC_HESS_CHUNK; k++) {
                    if (bb + k >= dim) continue;
                    hess[(ba + j) * dim + (bb + k)] = r.d[j].d[k];
                    hess[(bb + k) * dim + (ba + j)] = r.d[j].d[k];
                }
            }
        }
    }
    return value;
}

}  // namespace lmc
