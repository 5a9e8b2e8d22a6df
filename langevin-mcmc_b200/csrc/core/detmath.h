// detmath.h -- deterministic elementary functions shared by the sm_100a kernels and
// the host twin.
//
// Why: the parity target is an accept/reject sequence that is bit-identical between the
// GPU chain kernel and the CPU twin (SURVEY.md s7 "hard part 1").  glibc's libm and CUDA's
// libdevice round sinf/expf/... differently, so every transcendental on the hot path is
// re-stated here using only IEEE-754 +,-,*,/,sqrt,rint in double precision (correctly
// rounded on both x86-64 and sm_100a when contraction is off: `--fmad=false` for nvcc,
// `-ffp-contract=off` for gcc) and rounded once to float at the end.  Max error of each
// function is < 1 ulp(float); see tests/test_detmath.py.
//
// Reference call sites these replace: std::sin/cos (sampling.h:7-16, envlight.cpp:150-153),
// std::exp (microfacet.h:17, mutation_mala.h:264), std::log (microfacet.h:172,
// libstdc++ normal_distribution), std::pow (phong.cpp:46,110), acos/atan2
// (sampling.h:26-43, envlight.cpp:207-208), fastlog/fastpow (fastmath.h:364-381,1186-1190).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define LMC_HD __host__ __device__ __forceinline__
#define LMC_HD_NOINLINE __host__ __device__ __noinline__ inline
#else
#define LMC_HD inline
#define LMC_HD_NOINLINE inline
#endif

namespace lmc {

LMC_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LMC_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
LMC_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
LMC_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

LMC_HD bool dm_isfinite(float x) { return (f2u(x) & 0x7f800000u) != 0x7f800000u; }
LMC_HD bool dm_isnan(float x) { return x != x; }
LMC_HD float dm_inf() { return u2f(0x7f800000u); }
LMC_HD float dm_nan() { return u2f(0x7fc00000u); }

// NaN-transparent min/max with a fixed evaluation order (std::min/std::max semantics:
// min(a,b) = (b<a)?b:a ; max(a,b) = (a<b)?b:a).
LMC_HD float dm_min(float a, float b) { return (b < a) ? b : a; }
LMC_HD float dm_max(float a, float b) { return (a < b) ? b : a; }
LMC_HD float dm_clamp(float v, float lo, float hi) { return dm_min(dm_max(v, lo), hi); }
LMC_HD int dm_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
LMC_HD float dm_abs(float a) { return u2f(f2u(a) & 0x7fffffffu); }
LMC_HD float dm_sqrt(float a) { return sqrtf(a); }   // IEEE correctly rounded on both sides
LMC_HD float dm_floor(float a) { return floorf(a); } // exact
LMC_HD float dm_fmod(float a, float b) { return fmodf(a, b); }  // exact by definition

// ---------------------------------------------------------------------------------------
// double-precision kernels
// ---------------------------------------------------------------------------------------

// sin & cos of x (double), |x| < ~1e8.
LMC_HD_NOINLINE void dm_sincos_d(double x, double &s, double &c) {
    const double k = rint(x * 0.63661977236758134308);  // 2/pi
    double r = x - k * 1.57079632679489655800e+00;
    r = r - k * 6.12323399573676603587e-17;
    const double z = r * r;
    // Taylor, |r| <= pi/4
    double ps = -1.0 / 1307674368000.0;                  // r^15
    ps = ps * z + 1.0 / 6227020800.0;                    // r^13
    ps = ps * z - 1.0 / 39916800.0;                      // r^11
    ps = ps * z + 1.0 / 362880.0;                        // r^9
    ps = ps * z - 1.0 / 5040.0;                          // r^7
    ps = ps * z + 1.0 / 120.0;                           // r^5
    ps = ps * z - 1.0 / 6.0;                             // r^3
    const double sr = r + r * z * ps;
    double pc = 1.0 / 20922789888000.0;                  // r^16
    pc = pc * z - 1.0 / 87178291200.0;                   // r^14
    pc = pc * z + 1.0 / 479001600.0;                     // r^12
    pc = pc * z - 1.0 / 3628800.0;                       // r^10
    pc = pc * z + 1.0 / 40320.0;                         // r^8
    pc = pc * z - 1.0 / 720.0;                           // r^6
    pc = pc * z + 1.0 / 24.0;                            // r^4
    pc = pc * z - 0.5;                                   // r^2
    const double cr = 1.0 + z * pc;
    const int q = (int)((long long)k & 3LL);
    if (q == 0) { s = sr; c = cr; }
    else if (q == 1) { s = cr; c = -sr; }
    else if (q == 2) { s = -sr; c = -cr; }
    else { s = -cr; c = sr; }
}

// exp(x) for double x, result in double (clamped to float-representable magnitudes).
LMC_HD_NOINLINE double dm_exp_d(double x) {
    if (x != x) return x;
    if (x > 90.0) return (double)dm_inf();
    if (x < -110.0) return 0.0;
    const double k = rint(x * 1.44269504088896338700);
    double r = x - k * 6.93147180369123816490e-01;
    r = r - k * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;       // r^13
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    const long long e = (long long)k + 1023LL;   // in [864, 1153]
    return p * u2d((uint64_t)e << 52);
}

// log(x) for positive finite normal double x.
LMC_HD_NOINLINE double dm_log_pos_d(double x) {
    uint64_t b = d2u(x);
    long long e = (long long)((b >> 52) & 0x7ffULL) - 1023LL;
    double m = u2d((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.41421356237309514547) { m = m * 0.5; e = e + 1; }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    const double lm = 2.0 * s * p;
    return (double)e * 6.93147180559945286227e-01 + lm;
}

// atan(t) for 0 <= t <= 1
LMC_HD double dm_atan01_d(double t) {
    double base = 0.0;
    if (t > 0.41421356237309503) {
        t = (t - 1.0) / (t + 1.0);
        base = 0.78539816339744827900;
    }
    const double z = t * t;
    double p = -1.0 / 31.0;
    p = p * z + 1.0 / 29.0;
    p = p * z - 1.0 / 27.0;
    p = p * z + 1.0 / 25.0;
    p = p * z - 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z - 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z - 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z - 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z - 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z - 1.0 / 3.0;
    p = p * z + 1.0;
    return base + t * p;
}

LMC_HD_NOINLINE double dm_atan2_d(double y, double x) {
    const double ax = x < 0.0 ? -x : x;
    const double ay = y < 0.0 ? -y : y;
    if (ax == 0.0 && ay == 0.0) return 0.0;
    const double mn = ax < ay ? ax : ay;
    const double mx = ax < ay ? ay : ax;
    double a = dm_atan01_d(mn / mx);
    if (ay > ax) a = 1.57079632679489655800 - a;
    if (x < 0.0) a = 3.14159265358979311600 - a;
    if (y < 0.0) a = -a;
    return a;
}

// ---------------------------------------------------------------------------------------
// float front-ends
// ---------------------------------------------------------------------------------------
#if defined(LMC_TIMING_LIBM) && !defined(__CUDACC__)
// TIMING build of the CPU oracle only (`make oracle_fast`, used by bench.py's CPU arm): the platform libm in
// single precision, as the reference is built (`g++ -Ofast -march=native`, src/Tupfile:17).  Results are NOT
// bit-reproducible; no parity test uses this build.
LMC_HD float dm_sin(float x) { return sinf(x); }
LMC_HD float dm_cos(float x) { return cosf(x); }
LMC_HD void dm_sincos(float x, float &s, float &c) { s = sinf(x); c = cosf(x); }
LMC_HD float dm_exp(float x) { return expf(x); }
LMC_HD float dm_log(float x) { return logf(x); }
LMC_HD float dm_pow(float x, float y) { return powf(x, y); }
LMC_HD float dm_atan2(float y, float x) { return atan2f(y, x); }
LMC_HD float dm_acos(float x) { return acosf(x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x)); }
#else
LMC_HD float dm_sin(float x) { double s, c; dm_sincos_d((double)x, s, c); return (float)s; }
LMC_HD float dm_cos(float x) { double s, c; dm_sincos_d((double)x, s, c); return (float)c; }
LMC_HD void dm_sincos(float x, float &s, float &c) {
    double sd, cd; dm_sincos_d((double)x, sd, cd); s = (float)sd; c = (float)cd;
}
LMC_HD float dm_exp(float x) { return (float)dm_exp_d((double)x); }
LMC_HD float dm_log(float x) {
    if (x != x) return x;
    if (x < 0.0f) return dm_nan();
    if (x == 0.0f) return -dm_inf();
    if (!dm_isfinite(x)) return x;
    return (float)dm_log_pos_d((double)x);
}
// pow for x >= 0 (every reference call site has a non-negative base)
LMC_HD float dm_pow(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x != x || y != y) return dm_nan();
    if (x < 0.0f) return dm_nan();
    if (x == 0.0f) return y > 0.0f ? 0.0f : dm_inf();
    if (!dm_isfinite(x)) return y > 0.0f ? dm_inf() : 0.0f;
    return (float)dm_exp_d((double)y * dm_log_pos_d((double)x));
}
LMC_HD float dm_atan2(float y, float x) { return (float)dm_atan2_d((double)y, (double)x); }
// acos with the argument clamped to [-1,1] (the reference feeds normalised-vector components,
// sampling.h:38, envlight.cpp:208; libm would return NaN one ulp outside).
LMC_HD float dm_acos(float x) {
    double xd = (double)x;
    if (xd > 1.0) xd = 1.0;
    if (xd < -1.0) xd = -1.0;
    return (float)dm_atan2_d(sqrt((1.0 - xd) * (1.0 + xd)), xd);
}
#endif

// ---------------------------------------------------------------------------------------
// Mineiro fastlog / fastpow (bit tricks; fastmath.h:364-381, 233-242, 1186-1190).
// Re-stated from the published formulas; float ops only, no contraction.
// ---------------------------------------------------------------------------------------
LMC_HD float dm_fastlog2(float x) {
    const uint32_t vi = f2u(x);
    const float mf = u2f((vi & 0x007FFFFFu) | 0x3f000000u);
    float y = (float)vi;
    y *= 1.1920928955078125e-7f;
    return y - 124.22551499f - 1.498030302f * mf - 1.72587999f / (0.3520887068f + mf);
}
LMC_HD float dm_fastlog(float x) { return 0.69314718f * dm_fastlog2(x); }
LMC_HD float dm_fastpow2(float p) {
    const float offset = (p < 0.0f) ? 1.0f : 0.0f;
    const float clipp = (p < -126.0f) ? -126.0f : p;
    const int w = (int)clipp;
    const float z = clipp - (float)w + offset;
    const float v = (float)(1 << 23) *
                    (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z);
    return u2f((uint32_t)v);
}
LMC_HD float dm_fastpow(float x, float p) { return dm_fastpow2(p * dm_fastlog2(x)); }

}  // namespace lmc
