// trace_kernels.cuh -- the ray-query kernels of the wavefront: BVH2 closest-hit (k_trace) and any-hit
// (k_shadow) for whole queues of rays.  They replace Embree's rtcIntersect1 / rtcOccluded1
// (reference: src/scene.cpp:106-149) with the semantics pinned in core/bvh.h (same box_test / tri_test
// statements, closest hit = smallest t, ties -> smaller triangle id), so the answers are independent
// of the traversal order and bit-identical to the host twin's bvh_traverse.
//
// Execution model: PERSISTENT warps.  The rays of a queue are incoherent (paths of 2^20 independent
// chains) and need anything between 5 and 200 node visits, so a warp that takes 32 rays and runs until
// the last one finishes idles most of its lanes (measured 6-10 of 32 active).  Instead every warp keeps
// pulling rays from the queue cursor: whenever a quarter of its lanes have retired their ray they fetch
// new ones, and the while-while loop (descend to a leaf / intersect the leaf) keeps running on full
// warps (and leaves the descent loop early when only a few stragglers are still descending).  The top LMC_TOP_NODES nodes of the tree (breadth-first order, host_scene.cpp) are staged in
// shared memory with one TMA bulk copy per block (cp.async.bulk + mbarrier); deeper nodes and the
// triangles come through the read-only L1/L2 path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../core/bvh.h"

namespace lmc_cuda {
using namespace lmc;

#ifndef LMC_TOP_NODES
#define LMC_TOP_NODES 256          // 16 KB of shared memory per block (2 blocks per SM)
#endif
#ifndef LMC_TRACE_BLOCK
#define LMC_TRACE_BLOCK 512        // few, large blocks: shared memory taken from the L1 hurts (measured: 6 x 32 KB -> 2 x 16 KB = +4 %)
#endif
#ifndef LMC_TRACE_DESC_MIN
#define LMC_TRACE_DESC_MIN 12      // leave the descent loop when fewer lanes than this are still descending
#endif
#ifndef LMC_TRACE_REFILL
#define LMC_TRACE_REFILL 8         // refill when at least this many lanes are idle
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One elected thread arms an mbarrier with the byte count and issues the bulk copy global -> shared;
// everybody waits on the barrier's phase 0.
__device__ __forceinline__ void tma_stage_nodes(BvhNode *dst, const BvhNode *src, int count, uint64_t *bar) {
    const uint32_t barAddr = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)count * (uint32_t)sizeof(BvhNode);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(barAddr) : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LMC_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra LMC_TMA_DONE;\n"
        "bra LMC_TMA_WAIT;\n"
        "LMC_TMA_DONE:\n"
        "}\n" ::"r"(barAddr) : "memory");
}

struct F4s { float x, y, z, w; };
__device__ __forceinline__ F4s ld_node4(const BvhNode *top, int topCount, const BvhNode *nodes, int idx, int part) {
    F4s r;
    if (idx < topCount) {
        const float4 v = reinterpret_cast<const float4 *>(top + idx)[part];
        r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    } else {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(nodes + idx) + part);
        r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    }
    return r;
}

// SRC: int total(); bool load(int idx, Ray &ray, float &minT, float &maxT); void store(int idx, const Hit &h)
template <bool ANY_HIT, class SRC>
__device__ __forceinline__ void trace_persistent(const Scene &sc, const BvhNode *top, int topCount, SRC &src, int *cursor) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int total = src.total();
    int rayIdx = -1;
    bool exhausted = false;          // warp-uniform: the queue has no rays left to hand out
    Ray ray; ray.org = mk3s(0.0f); ray.dir = mk3s(0.0f);
    V3 invDir = mk3s(0.0f), negOrgInv = mk3s(0.0f);
    float minT = 0.0f;
    Hit best; best.tid = -1; best.t = 0.0f; best.u = 0.0f; best.v = 0.0f;
    int stack[LMC_BVH_STACK];
    int sp = 0, cur = LMC_BVH_DONE;
    for (;;) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(FULL, rayIdx < 0);
        if (!exhausted && (__popc(idle) >= LMC_TRACE_REFILL || idle == FULL)) {
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, __popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (base + __popc(idle) >= total) exhausted = true;
            if (rayIdx < 0) {
                const int idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < total) {
                    float maxT;
                    src.load(idx, ray, minT, maxT);
                    rayIdx = idx;
                    best.tid = -1; best.t = maxT; best.u = 0.0f; best.v = 0.0f;
                    invDir = mk3(inverse(ray.dir.x), inverse(ray.dir.y), inverse(ray.dir.z));
                    negOrgInv = mk3(-(ray.org.x * invDir.x), -(ray.org.y * invDir.y), -(ray.org.z * invDir.z));
                    sp = 0; cur = (sc.numNodes > 0) ? 0 : LMC_BVH_DONE;
                }
            }
        }
        if (__ballot_sync(FULL, rayIdx >= 0) == 0u) break;
        // ---- descend to the next leaf.  The number of node visits until a lane reaches its next leaf is
        // heavy-tailed (measured: 7 of 32 lanes active in a plain `while (cur >= 0)` loop), so the warp leaves
        // the loop as soon as fewer than LMC_TRACE_DESC_MIN lanes are still descending while others hold a
        // leaf; the stragglers simply keep descending in the next turn, next to the lanes that come back
        // from their leaves.
        for (;;) {
            const bool desc = cur >= 0;
            const unsigned dm = __ballot_sync(FULL, desc);
            if (dm == 0u) break;
            if (__popc(dm) < LMC_TRACE_DESC_MIN && __ballot_sync(FULL, cur < 0 && cur != LMC_BVH_DONE) != 0u) break;
            if (desc) {
                const F4s a = ld_node4(top, topCount, sc.nodes, cur, 0);     // lmin.xyz, lmax.x
                const F4s b = ld_node4(top, topCount, sc.nodes, cur, 1);     // lmax.yz, rmin.xy
                const F4s c = ld_node4(top, topCount, sc.nodes, cur, 2);     // rmin.z, rmax.xyz
                const F4s d = ld_node4(top, topCount, sc.nodes, cur, 3);     // left, right
                float tl, tr;
                const bool hl = box_test(a.x, a.y, a.z, a.w, b.x, b.y, invDir, negOrgInv, minT, best.t, tl);
                const bool hr = box_test(b.z, b.w, c.x, c.y, c.z, c.w, invDir, negOrgInv, minT, best.t, tr);
                const int left = __float_as_int(d.x), right = __float_as_int(d.y);
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
        }
        // ---- intersect the leaf
        if (cur < 0 && cur != LMC_BVH_DONE) {
            const int enc = ~cur;
            const int first = enc >> 3;
            const int count = (enc & 7) + 1;
            bool found = false;
            for (int i = 0; i < count; ++i) {
                const int tid = first + i;
                float t, u, v;
                if (tri_test(sc.tris[tid], ray, minT, best.t, t, u, v)) {
                    if (ANY_HIT) { best.tid = tid; best.t = t; best.u = u; best.v = v; found = true; break; }
                    if (t < best.t || best.tid < 0 || tid < best.tid) { best.tid = tid; best.t = t; best.u = u; best.v = v; }
                }
            }
            cur = (ANY_HIT && found) ? LMC_BVH_DONE : ((sp > 0) ? stack[--sp] : LMC_BVH_DONE);
        }
        // ---- retire finished rays
        if (rayIdx >= 0 && cur == LMC_BVH_DONE) {
            src.store(rayIdx, best);
            rayIdx = -1;
        }
    }
}

}  // namespace lmc_cuda
