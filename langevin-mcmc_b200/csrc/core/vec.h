// vec.h -- minimal 2/3-vector and 4x4 matrix helpers with a fixed, documented operation order
// (left-to-right sums, multiply-by-reciprocal normalisation) so host twin and device agree
// bit for bit.  Mirrors the helper semantics of the reference's src/utils.h:110-232 and
// src/transform.h:46-72 (Dot, Cross, Normalize = v * inverse(Length(v)), XformPoint with
// homogeneous divide, XformVector).
#pragma once
#include "detmath.h"

namespace lmc {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct M44 { float m[4][4]; };  // row-major m[r][c]

LMC_HD V2 mk2(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
LMC_HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
LMC_HD V3 mk3s(float s) { return mk3(s, s, s); }

LMC_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
LMC_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
LMC_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
LMC_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
LMC_HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
LMC_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
LMC_HD V3 cmul(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
LMC_HD V3 &operator+=(V3 &a, V3 b) { a = a + b; return a; }
LMC_HD V3 &operator*=(V3 &a, float s) { a = a * s; return a; }

LMC_HD float inverse(float x) { return 1.0f / x; }
LMC_HD float square(float x) { return x * x; }
// Fused multiply-adds are written out explicitly (fmaf is correctly rounded on x86-64 and on
// sm_100a), so host twin and device agree bit for bit while the device issues FFMA; compiler
// contraction stays off on both sides.
LMC_HD float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
LMC_HD V3 cross(V3 a, V3 b) {
    return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
LMC_HD float length_squared(V3 v) { return fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)); }
LMC_HD float length(V3 v) { return dm_sqrt(fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x))); }
LMC_HD float distance_squared(V3 a, V3 b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
// a * s + b
LMC_HD V3 madd(V3 a, float s, V3 b) { return mk3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }
LMC_HD V3 normalize(V3 v) { const float il = inverse(length(v)); return v * il; }
LMC_HD float luminance(V3 v) { return fmaf(v.z, 0.072169f, fmaf(v.y, 0.715160f, v.x * 0.212671f)); }
LMC_HD bool is_zero(V3 v) { return v.x == 0.0f && v.y == 0.0f && v.z == 0.0f; }
LMC_HD float max_coeff(V3 v) { return dm_max(dm_max(v.x, v.y), v.z); }
LMC_HD bool all_finite(V3 v) { return dm_isfinite(v.x) && dm_isfinite(v.y) && dm_isfinite(v.z); }
LMC_HD float comp(const V3 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// src/utils.h:198-201
LMC_HD V3 reflect(V3 wi, V3 n) { return madd(n, 2.0f * dot(wi, n), -wi); }
// src/utils.h:203-210
LMC_HD V3 refract(V3 wi, V3 n, float cosThetaT, float eta, float invEta) {
    const float eta_ = (cosThetaT < 0.0f) ? invEta : eta;
    return n * (dot(wi, n) * eta_ + cosThetaT) - wi * eta_;
}
// src/utils.h:222-232
LMC_HD void coordinate_system(V3 n, V3 &b1, V3 &b2) {
    if (n.z < (float)(-1.0 + 1e-6)) {
        b1 = mk3(0.0f, -1.0f, 0.0f);
        b2 = mk3(-1.0f, 0.0f, 0.0f);
        return;
    }
    const float a = 1.0f / (1.0f + n.z);
    const float b = -n.x * n.y * a;
    b1 = mk3(1.0f - square(n.x) * a, b, -n.x);
    b2 = mk3(b, 1.0f - square(n.y) * a, -n.y);
}

// src/transform.h:46-57
LMC_HD V3 xform_point(const M44 &t, V3 p) {
    const float x = fmaf(t.m[0][2], p.z, fmaf(t.m[0][1], p.y, fmaf(t.m[0][0], p.x, t.m[0][3])));
    const float y = fmaf(t.m[1][2], p.z, fmaf(t.m[1][1], p.y, fmaf(t.m[1][0], p.x, t.m[1][3])));
    const float z = fmaf(t.m[2][2], p.z, fmaf(t.m[2][1], p.y, fmaf(t.m[2][0], p.x, t.m[2][3])));
    const float w = fmaf(t.m[3][2], p.z, fmaf(t.m[3][1], p.y, fmaf(t.m[3][0], p.x, t.m[3][3])));
    const float iw = inverse(w);
    return mk3(x * iw, y * iw, z * iw);
}
// src/transform.h:59-65
LMC_HD V3 xform_vector(const M44 &t, V3 v) {
    return mk3(fmaf(t.m[0][2], v.z, fmaf(t.m[0][1], v.y, t.m[0][0] * v.x)),
               fmaf(t.m[1][2], v.z, fmaf(t.m[1][1], v.y, t.m[1][0] * v.x)),
               fmaf(t.m[2][2], v.z, fmaf(t.m[2][1], v.y, t.m[2][0] * v.x)));
}
// transpose(M) * v  (adjoint of xform_vector)
LMC_HD V3 xform_vector_t(const M44 &t, V3 v) {
    return mk3(t.m[0][0] * v.x + t.m[1][0] * v.y + t.m[2][0] * v.z,
               t.m[0][1] * v.x + t.m[1][1] * v.y + t.m[2][1] * v.z,
               t.m[0][2] * v.x + t.m[1][2] * v.y + t.m[2][2] * v.z);
}

// src/utils.h:382-385 (Modulo for Float)
LMC_HD float modulo1(float a) {
    const float r = dm_fmod(a, 1.0f);
    return (r < 0.0f) ? r + 1.0f : r;
}

}  // namespace lmc
