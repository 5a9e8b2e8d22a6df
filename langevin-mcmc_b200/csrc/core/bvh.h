// bvh.h -- BVH2 closest-hit / any-hit traversal replacing Embree's rtcIntersect1 /
// rtcOccluded1 (reference: src/scene.cpp:106-149) and the per-triangle re-intersection
// TriangleMesh::Intersect (src/trianglemesh.cpp:30-79,189-236).
//
// Semantics pinned here (Embree's own tie-breaking is unpinned):
//   * leaf test = the reference's Moeller-Trumbore formulas (same operation order), hit iff
//     divisor != 0, u >= 0, v >= 0, u + v <= 1, minT <= t <= maxT;
//   * closest hit = smallest t, ties broken by the smaller triangle id (BVH order), so the
//     answer does not depend on traversal order;
//   * any hit = first triangle found with a hit inside [minT, maxT].
#pragma once
#include "scene.h"

namespace lmc {

struct Ray { V3 org, dir; };

struct Hit {
    int tid;       // triangle id (BVH order) or -1
    float t, u, v;
};

// 16-byte loads (read-only data path on the device)
struct F4 { float x, y, z, w; };
LMC_HD F4 ld4(const void *p) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    F4 r; memcpy(&r, p, 16); return r;
#endif
}
// min / max that lower to single FMNMX instructions; IEEE fmin/fmax semantics on both sides
LMC_HD float bmin(float a, float b) { return fminf(a, b); }
LMC_HD float bmax(float a, float b) { return fmaxf(a, b); }

LMC_HD bool tri_test(const TriGeom &tg, const Ray &ray, float minT, float maxT,
                     float &t, float &u, float &v) {
    const F4 q0 = ld4(tg.p0), q1 = ld4(tg.e1), q2 = ld4(tg.e2);
    const V3 p0 = mk3(q0.x, q0.y, q0.z), e1 = mk3(q1.x, q1.y, q1.z), e2 = mk3(q2.x, q2.y, q2.z);
    const V3 s1 = cross(ray.dir, e2);
    const float divisor = dot(s1, e1);
    if (divisor == 0.0f) return false;
    const float invDivisor = inverse(divisor);
    const V3 s = ray.org - p0;
    u = dot(s, s1) * invDivisor;
    const V3 s2 = cross(s, e1);
    v = dot(ray.dir, s2) * invDivisor;
    if (!(u >= 0.0f && v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, s2) * invDivisor;
    return (t >= minT && t <= maxT);
}

// Slab test of one child box.  t = b * invDir - org * invDir evaluated with ONE fused
// multiply-add per plane (fmaf is correctly rounded on x86-64 and sm_100a alike, so the host twin
// takes the same traversal decisions).  Conservative culling only: hit results always come from
// tri_test.
LMC_HD bool box_test(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, const V3 &invDir,
                     const V3 &negOrgInv, float minT, float maxT, float &tNear) {
    float t0 = fmaf(bminx, invDir.x, negOrgInv.x), t1 = fmaf(bmaxx, invDir.x, negOrgInv.x);
    float lo = bmin(t0, t1), hi = bmax(t0, t1);
    t0 = fmaf(bminy, invDir.y, negOrgInv.y); t1 = fmaf(bmaxy, invDir.y, negOrgInv.y);
    lo = bmax(lo, bmin(t0, t1)); hi = bmin(hi, bmax(t0, t1));
    t0 = fmaf(bminz, invDir.z, negOrgInv.z); t1 = fmaf(bmaxz, invDir.z, negOrgInv.z);
    lo = bmax(lo, bmin(t0, t1)); hi = bmin(hi, bmax(t0, t1));
    lo = bmax(lo, minT); hi = bmin(hi, maxT);
    tNear = lo;
    return lo <= hi;
}

#define LMC_BVH_STACK 32
#define LMC_BVH_DONE ((int)0x80000000)

// "while-while" traversal: every lane first descends through inner nodes until it holds a leaf
// (or runs out of work), then all lanes intersect their leaves together -- the two loop bodies no
// longer serialise against each other inside a warp.
#ifdef LMC_BVH_STATS
static long long g_bvhNodes = 0, g_bvhTris = 0, g_bvhRays = 0;
#endif
template <bool ANY_HIT>
LMC_HD_NOINLINE Hit bvh_traverse(const Scene &sc, const Ray &ray, float minT, float maxT) {
    Hit best; best.tid = -1; best.t = maxT; best.u = 0.0f; best.v = 0.0f;
    if (sc.numNodes == 0) return best;
#ifdef LMC_BVH_STATS
    g_bvhRays++;
#endif
    const V3 invDir = mk3(inverse(ray.dir.x), inverse(ray.dir.y), inverse(ray.dir.z));
    const V3 negOrgInv = mk3(-(ray.org.x * invDir.x), -(ray.org.y * invDir.y), -(ray.org.z * invDir.z));
    int stack[LMC_BVH_STACK];
    int sp = 0;
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
#ifdef LMC_BVH_STATS
            g_bvhNodes++;
#endif
            const BvhNode *n = sc.nodes + cur;
            const F4 a = ld4(&n->lmin[0]);     // lmin.xyz, lmax.x
            const F4 b = ld4(&n->lmax[1]);     // lmax.yz, rmin.xy
            const F4 c = ld4(&n->rmin[2]);     // rmin.z, rmax.xyz
            const F4 d = ld4(&n->left);        // left, right, pad
            float tl, tr;
            const bool hl = box_test(a.x, a.y, a.z, a.w, b.x, b.y, invDir, negOrgInv, minT, best.t, tl);
            const bool hr = box_test(b.z, b.w, c.x, c.y, c.z, c.w, invDir, negOrgInv, minT, best.t, tr);
            const int left = (int)f2u(d.x), right = (int)f2u(d.y);
            if (hl && hr) {
                const bool swap = tr < tl;
                const int nearC = swap ? right : left, farC = swap ? left : right;
                if (sp < LMC_BVH_STACK) stack[sp++] = farC;
                cur = nearC;
            } else if (hl) {
                cur = left;
            } else if (hr) {
                cur = right;
            } else {
                cur = (sp > 0) ? stack[--sp] : LMC_BVH_DONE;
            }
        }
        if (cur == LMC_BVH_DONE) break;
        {
            const int enc = ~cur;
            const int first = enc >> 3;
            const int count = (enc & 7) + 1;
            for (int i = 0; i < count; ++i) {
#ifdef LMC_BVH_STATS
                g_bvhTris++;
#endif
                const int tid = first + i;
                float t, u, v;
                if (tri_test(sc.tris[tid], ray, minT, best.t, t, u, v)) {
                    if (ANY_HIT) { best.tid = tid; best.t = t; best.u = u; best.v = v; return best; }
                    if (t < best.t || best.tid < 0 || tid < best.tid) {
                        best.tid = tid; best.t = t; best.u = u; best.v = v;
                    }
                }
            }
        }
        cur = (sp > 0) ? stack[--sp] : LMC_BVH_DONE;
    }
    return best;
}

// Brute-force closest hit with identical semantics (test/oracle use only).
LMC_HD Hit brute_closest(const Scene &sc, const Ray &ray, float minT, float maxT) {
    Hit best; best.tid = -1; best.t = maxT; best.u = 0.0f; best.v = 0.0f;
    for (int tid = 0; tid < sc.numTris; ++tid) {
        float t, u, v;
        if (tri_test(sc.tris[tid], ray, minT, best.t, t, u, v)) {
            if (t < best.t || best.tid < 0) { best.tid = tid; best.t = t; best.u = u; best.v = v; }
        }
    }
    return best;
}

struct Isect {
    V3 position, shadingNormal, geomNormal;
};

// Fill the intersection record for triangle `tid` hit at (t,u,v)
// (reference: TriangleIntersect + TriangleMesh::Intersect, src/trianglemesh.cpp:58-79,189-236).
LMC_HD_NOINLINE void fill_isect(const Scene &sc, const Ray &ray, const Hit &h, Isect &isect, V2 &st) {
    const TriGeom &tg = sc.tris[h.tid];
    const TriShade &ts = sc.shade[h.tid];
    const V3 e1 = ld3(tg.e1), e2 = ld3(tg.e2);
    isect.geomNormal = normalize(cross(e1, e2));
    const float w = 1.0f - h.u - h.v;
    isect.position = madd(ray.dir, h.t, ray.org);
    isect.shadingNormal = normalize(w * ld3(ts.n0) + h.u * ld3(ts.n1) + h.v * ld3(ts.n2));
    if (dot(isect.geomNormal, isect.shadingNormal) < 0.0f) {
        isect.geomNormal = -isect.geomNormal;
    }
    if (sc.mats[tg.geom].hasST) {
        st.x = (1.0f - h.u - h.v) * ts.st0[0] + h.u * ts.st1[0] + h.v * ts.st2[0];
        st.y = (1.0f - h.u - h.v) * ts.st0[1] + h.u * ts.st1[1] + h.v * ts.st2[1];
    } else {
        st = mk2(h.u, h.v);
    }
}


// Intersect(scene, time, raySeg, shapeInst, isect)  (src/path.cpp:90-102)
LMC_HD bool scene_intersect(const Scene &sc, const Ray &ray, float minT, float maxT,
                            int &tid, Isect &isect, V2 &st) {
    const Hit h = bvh_traverse<false>(sc, ray, minT, maxT);
    if (h.tid < 0) return false;
    tid = h.tid;
    fill_isect(sc, ray, h, isect, st);
    return true;
}

// Occluded(scene, time, ray, dist)  (src/scene.cpp:128-149)
LMC_HD bool scene_occluded(const Scene &sc, const Ray &ray, float dist) {
    const float minT = LMC_ISECT_EPS;
    float maxT;
    if (dist == dm_inf()) maxT = dm_inf();
    else maxT = (1.0f - LMC_SHADOW_EPS) * dist;
    const Hit h = bvh_traverse<true>(sc, ray, minT, maxT);
    return h.tid >= 0;
}

}  // namespace lmc
