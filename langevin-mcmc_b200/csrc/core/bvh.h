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

LMC_HD bool tri_test(const TriGeom &tg, const Ray &ray, float minT, float maxT,
                     float &t, float &u, float &v) {
    const V3 p0 = ld3(tg.p0), e1 = ld3(tg.e1), e2 = ld3(tg.e2);
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

LMC_HD bool box_test(const float *bmin, const float *bmax, const V3 &org, const V3 &invDir,
                     float minT, float maxT, float &tNear) {
    float t0 = (bmin[0] - org.x) * invDir.x, t1 = (bmax[0] - org.x) * invDir.x;
    float lo = dm_min(t0, t1), hi = dm_max(t0, t1);
    t0 = (bmin[1] - org.y) * invDir.y; t1 = (bmax[1] - org.y) * invDir.y;
    lo = dm_max(lo, dm_min(t0, t1)); hi = dm_min(hi, dm_max(t0, t1));
    t0 = (bmin[2] - org.z) * invDir.z; t1 = (bmax[2] - org.z) * invDir.z;
    lo = dm_max(lo, dm_min(t0, t1)); hi = dm_min(hi, dm_max(t0, t1));
    lo = dm_max(lo, minT); hi = dm_min(hi, maxT);
    tNear = lo;
    return lo <= hi;
}

#define LMC_BVH_STACK 48

template <bool ANY_HIT>
LMC_HD_NOINLINE Hit bvh_traverse(const Scene &sc, const Ray &ray, float minT, float maxT) {
    Hit best; best.tid = -1; best.t = maxT; best.u = 0.0f; best.v = 0.0f;
    if (sc.numNodes == 0) return best;
    const V3 invDir = mk3(inverse(ray.dir.x), inverse(ray.dir.y), inverse(ray.dir.z));
    int stack[LMC_BVH_STACK];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const BvhNode &n = sc.nodes[cur];
            float tl, tr;
            const bool hl = box_test(n.lmin, n.lmax, ray.org, invDir, minT, best.t, tl);
            const bool hr = box_test(n.rmin, n.rmax, ray.org, invDir, minT, best.t, tr);
            if (hl && hr) {
                int nearC = n.left, farC = n.right;
                if (tr < tl) { nearC = n.right; farC = n.left; }
                if (sp < LMC_BVH_STACK) stack[sp++] = farC;
                cur = nearC;
                continue;
            } else if (hl) {
                cur = n.left; continue;
            } else if (hr) {
                cur = n.right; continue;
            }
        } else {
            const int enc = ~cur;
            const int first = enc >> 3;
            const int count = (enc & 7) + 1;
            for (int i = 0; i < count; ++i) {
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
        if (sp == 0) break;
        cur = stack[--sp];
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
    isect.position = ray.org + h.t * ray.dir;
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
