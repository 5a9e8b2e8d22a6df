// stages.h -- resumable ("staged") forms of PerturbPathBidir and GeneratePathBidir for the
// per-vertex wavefront: every function here runs the reference's statements between two
// closest-hit ray queries, so the device can do ALL ray queries of an iteration in dedicated
// traversal kernels (cuda/trace_kernels.cuh) and everything else in small shading kernels.
//
//   PerturbPathBidir   src/path.cpp:1953-2160   ->  perturb_stage_begin / _light / _camera
//   GeneratePathBidir  src/path.cpp:1237-1449   ->  gen_stage_begin / _light / _camera
//
// The statements, their order and the order of the RNG draws are those of path.h's monolithic
// functions (which stay: the host twin calls them, and tests/test_staged.py checks that both
// forms produce bit-identical chains).  Two things differ in mechanism, not in result:
//   * a closest-hit query is "return true with ts.ray/minT/maxT set"; the caller answers with a
//     Hit (tid, t, u, v) and calls the next stage function;
//   * visibility of connection segments is DEFERRED: ConnectToCamera / DirectLighting /
//     ConnectVertex evaluate the whole contribution first and ask for visibility last
//     (path.h), DeferredList answers "visible", records the contribution as PENDING and hands
//     the shadow ray to a sink; when the ray has been traced the flag becomes VISIBLE or
//     OCCLUDED, and deferred_compact() rebuilds the reference's contribution vector in order.
#pragma once
#include "path.h"

namespace lmc {

enum TraceStage { TS_DONE = 0, TS_P_LGT = 1, TS_P_CAM = 2, TS_G_LGT = 3, TS_G_CAM = 4 };
enum CandFlag { CAND_VISIBLE = 0, CAND_PENDING = 1, CAND_OCCLUDED = 2, CAND_CLEAR = 4 };

// Capacity invariant.  The list holds, per path function call, the contributions the reference's vector would hold
// (bounded by Limits<MAXD>::MAXC - 2, mutation.h), the slots reserved for deferred connections (they ARE such
// contributions) and at most ONE clear entry: every `contribs.clear()` of the path functions is followed by `return`
// (src/path.cpp:704-708,1384-1387,2075-2078).  So n <= MAXC - 1 and the `n < cap` guards below never fire; if the bound
// were ever wrong an entry would be dropped (and an unconditional clear would fall back to rewinding the list), which the
// host build counts -- the staged-path tests assert the counter stays 0 (lmco_deferred_overflows).
#if !defined(__CUDACC__)
inline long &deferred_overflow_count() { static long c = 0; return c; }
#define LMC_DEFERRED_OVERFLOW() (++deferred_overflow_count())
#else
#define LMC_DEFERRED_OVERFLOW() ((void)0)
#endif

// SINK: void emit(const Ray &ray, float dist, int slot, int *flag)
template <class SINK>
struct DeferredList {
    SubpathContrib *c;
    int *flag;
    int *np;            // number of entries (lives with the chain)
    int cap;
    SINK *sink;
    bool pend;
    Ray pray;
    float pdist;
    LMC_HD void bind(SubpathContrib *c_, int *flag_, int *np_, int cap_, SINK *sink_) {
        c = c_; flag = flag_; np = np_; cap = cap_; sink = sink_; pend = false;
    }
    LMC_HD bool occluded(const Scene &, const Ray &ray, float dist) { pend = true; pray = ray; pdist = dist; return false; }
    LMC_HD void push(const SubpathContrib &x) {
        const int n = *np;
        if (n < cap) {
            c[n] = x;
            flag[n] = pend ? CAND_PENDING : CAND_VISIBLE;
            *np = n + 1;
            if (pend) sink->emit(pray, pdist, n, flag + n);
        } else LMC_DEFERRED_OVERFLOW();
        pend = false;
    }
    // ConnectVertex of GeneratePathBidir: evaluated on the spot, or -- when the sink wants it
    // (SINK::kDeferConnections) -- a slot is reserved in push order (flagged OCCLUDED = "nothing here" until
    // somebody fills it) and the pair is handed to the sink, which evaluates it later into that slot.
    template <class LS>
    LMC_HD void connect(const Scene &sc, int camDepth, int lgtDepth, const LS *ls, const SurfaceVertex *lgtVerts,
                        const LS &cps, const SurfaceVertex &camVertex, V2 screenPos) {
        if (SINK::kDeferConnections) {
            const int n = *np;
            if (n < cap) {
                flag[n] = CAND_OCCLUDED;
                *np = n + 1;
                sink->emit_connection(sc, camDepth, lgtDepth, n, ls, lgtVerts, cps, camVertex, screenPos, c + n, flag + n);
            } else LMC_DEFERRED_OVERFLOW();
        } else {
            const SurfaceVertex lv = lgtVerts[lgtDepth];
            connect_vertex(sc, camDepth, lgtDepth, ls[lgtDepth], lv, cps, camVertex, screenPos, *this);
        }
    }
    // clear() right after a visibility query = "clear if that segment is visible"
    LMC_HD void clear() {
        if (pend) {
            const int n = *np;
            if (n < cap) {
                flag[n] = CAND_PENDING | CAND_CLEAR;
                *np = n + 1;
                sink->emit(pray, pdist, n, flag + n);
            } else LMC_DEFERRED_OVERFLOW();
            pend = false;
        } else {
            // unconditional clear: recorded as an already-visible CLEAR entry rather than by rewinding the
            // list, so that slots whose shadow rays are still in flight are never reused
            const int n = *np;
            if (n < cap) { flag[n] = CAND_VISIBLE | CAND_CLEAR; *np = n + 1; }
            else { *np = 0; LMC_DEFERRED_OVERFLOW(); }
        }
    }
};

// Contribution sink bound to ONE reserved slot (a deferred ConnectVertex evaluates into it)
template <class SINK>
struct SlotList {
    SubpathContrib *c;
    int *flag;
    SINK *sink;
    bool pend;
    Ray pray;
    float pdist;
    LMC_HD bool occluded(const Scene &, const Ray &ray, float dist) { pend = true; pray = ray; pdist = dist; return false; }
    LMC_HD void push(const SubpathContrib &x) {
        *c = x;
        *flag = pend ? CAND_PENDING : CAND_VISIBLE;
        if (pend) sink->emit(pray, pdist, 0, flag);
        pend = false;
    }
    LMC_HD void clear() {}      // ConnectVertex never clears
};

// what the shadow-ray kernel does with its answer
LMC_HD int cand_resolve(int flag, bool occluded) { return (flag & CAND_CLEAR) | (occluded ? CAND_OCCLUDED : CAND_VISIBLE); }

// Rebuild the reference's contribution vector (in push order) from resolved candidates, in place.
LMC_HD int deferred_compact(SubpathContrib *c, const int *flag, int n) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        const int f = flag[i];
        if (f & CAND_OCCLUDED) continue;
        if (f & CAND_CLEAR) { m = 0; continue; }
        if (m != i) c[m] = c[i];
        m++;
    }
    return m;
}

// State of a proposal between two stages.  The device moves it from ray queue to ray queue as
// the wavefront payload (together with the PathHead), 16 bytes at a time: keep it a multiple of 16 B
// and keep the ray at words 6..13 (cuda/chain_kernels.cuh k_trace reads it from there).
struct alignas(16) TraceState {
    int stage;            // TraceStage: which function consumes the next Hit
    int depth;            // lgtDepth / camDepth of the vertex the pending ray will find
    int offsetId;
    int nLightStates;
    float ndSaved;        // NormalDist(0, discreteStdDev) carry of PerturbPathBidir
    int ndAvail;
    float minT, maxT;
    Ray ray;
    uint32_t rngLo, rngHi;   // the chain's RNG state rides along (device); unused by the host model
    uint32_t rngEpoch;
    int curIdx;              // which MarkovState slot is the current state (device)
    int tsPad[2];
    BidirPathState cps, lps;
    int tsPad2[2];
};

// BidirPathState padded to 16 bytes (the device parks the light-subpath state of a perturbation in it)
struct alignas(16) LpsFull { BidirPathState s; float pad; };

// Candidate contribution(s) of a small step: at most one (+ a conditional clear)
struct PropCand {
    int n;
    int flag[2];
    SubpathContrib c[2];
};

// Camera-subpath vertex as a deferred ConnectVertex needs it (device: written once per camera vertex)
struct alignas(16) CamSnap {
    BidirPathState cps;
    int tid; V2 st;          // the camera vertex (shapeInst, st)
    V2 screenPos;
    int curIdx;
    int pad[3];
};

// Large-step workspace: light subpath states and the contribution candidates.
template <int MAXD, int MAXC>
struct GenWork {
    CamSnap snap[MAXD];
    BidirPathState ls[MAXD];
    int n;
    int flag[MAXC];
    SubpathContrib c[MAXC];
};

LMC_HD void trace_state_init(TraceState &ts) { memset(&ts, 0, sizeof(ts)); }

// ---- PerturbPathBidir --------------------------------------------------------------------
// `ph` is the proposal's PathHead, `sv` the vertex the pending ray was shot for (path.lgt[ts.depth] /
// path.cam[ts.depth]), `lgtVerts` = path.lgt.  OFF: anything indexable like the offset vector.
template <class OFF>
LMC_HD bool perturb_camera_begin(const Scene &sc, const OFF &offset, PathHead &ph, TraceState &ts) {
    perturb(ph.screenPos.x, offset, ts.offsetId);
    perturb(ph.screenPos.y, offset, ts.offsetId);
    emit_from_camera(sc, ph.screenPos, ts.ray, ts.minT, ts.maxT, ts.cps);
    if (ph.nCam <= 0) { ts.stage = TS_DONE; return false; }
    ts.stage = TS_P_CAM; ts.depth = 0;
    return true;
}

template <class OFF>
LMC_HD bool perturb_stage_begin(const Scene &sc, const OFF &offset, PathHead &ph, TraceState &ts, Rng &rng) {
    NormalDist nd = normal_make(0.0f, sc.opt.discreteStdDev);
    ts.offsetId = 0;
    ph.time = modulo1(ph.time + normal_draw(nd, rng));
    ts.ndSaved = nd.saved; ts.ndAvail = nd.savedAvailable ? 1 : 0;
    if (ph.lgtDepth > 1) {
        const float lightPickProb = pick_light_prob(sc, ph.lgtLight);
        perturb(ph.lgtRndPos.x, offset, ts.offsetId);
        perturb(ph.lgtRndPos.y, offset, ts.offsetId);
        perturb(ph.lgtRndDir.x, offset, ts.offsetId);
        perturb(ph.lgtRndDir.y, offset, ts.offsetId);
        emit_from_light(sc, lightPickProb, ph, ts.ray, ts.lps);
        ts.minT = LMC_ISECT_EPS; ts.maxT = dm_inf();
        if (ph.nLgt > 0) { ts.stage = TS_P_LGT; ts.depth = 0; return true; }
    }
    return perturb_camera_begin(sc, offset, ph, ts);
}

template <class OFF, class CL>
LMC_HD bool perturb_stage_light(const Scene &sc, const OFF &offset, PathHead &ph, SurfaceVertex &sv, TraceState &ts,
                                CL &contribs, Rng &rng, const Hit &hit) {
    const int lgtDepth = ts.depth;
    BidirPathState &lps = ts.lps;
    if (hit.tid < 0) { ts.stage = TS_DONE; return false; }
    sv.tid = hit.tid;
    fill_isect(sc, ts.ray, hit, lps.isect, sv.st);
    lps.wi = -ts.ray.dir;
    NormalDist nd = normal_make(0.0f, sc.opt.discreteStdDev);
    nd.saved = ts.ndSaved; nd.savedAvailable = ts.ndAvail != 0;
    sv.bsdfDiscrete = modulo1(sv.bsdfDiscrete + normal_draw(nd, rng));
    ts.ndSaved = nd.saved; ts.ndAvail = nd.savedAvailable ? 1 : 0;
    convert_mis(sc, lgtDepth, ph.lgtLight, ts.ray, lps);
    if (lgtDepth == ph.nLgt - 1 && ph.camDepth == 1) {
        connect_to_camera(sc, lgtDepth, lps, sv, ts.ray.org, contribs);
        ts.stage = TS_DONE; return false;
    }
    if (lgtDepth == ph.nLgt - 1) return perturb_camera_begin(sc, offset, ph, ts);
    perturb(sv.bsdfRndParam.x, offset, ts.offsetId);
    perturb(sv.bsdfRndParam.y, offset, ts.offsetId);
    V3 bsdfContrib;
    if (!bsdf_sampling<true, true>(sc, lps, sv, lps, ts.ray.dir, bsdfContrib)) { ts.stage = TS_DONE; return false; }
    lps.throughput *= sv.rrWeight;
    ts.ray.org = lps.isect.position;
    ts.depth = lgtDepth + 1;
    return true;
}

template <class OFF, class CL>
LMC_HD bool perturb_stage_camera(const Scene &sc, const OFF &offset, PathHead &ph, SurfaceVertex &sv,
                                 const SurfaceVertex *lgtVerts, TraceState &ts, CL &contribs, Rng &rng, const Hit &hit) {
    const int camDepth = ts.depth;
    BidirPathState &cps = ts.cps;
    const bool hitSurface = hit.tid >= 0;
    if (hitSurface) { sv.tid = hit.tid; fill_isect(sc, ts.ray, hit, cps.isect, sv.st); }
    cps.wi = -ts.ray.dir;
    if (hitSurface) convert_mis(sc, camDepth, -1, ts.ray, cps);
    if (camDepth == ph.nCam - 1 && ph.lgtDepth == 0) {
        const int light = get_hit_light(sc, hitSurface, sv.tid);
        if (light >= 0) handle_hit_light(sc, camDepth, light, hitSurface, ts.ray, ph.screenPos, cps, ph, contribs);
        ts.stage = TS_DONE; return false;
    }
    if (!hitSurface) { ts.stage = TS_DONE; return false; }
    NormalDist nd = normal_make(0.0f, sc.opt.discreteStdDev);
    nd.saved = ts.ndSaved; nd.savedAvailable = ts.ndAvail != 0;
    sv.bsdfDiscrete = modulo1(sv.bsdfDiscrete + normal_draw(nd, rng));
    ts.ndSaved = nd.saved; ts.ndAvail = nd.savedAvailable ? 1 : 0;
    if (camDepth == 1) {
        ph.lensVertexPos = cps.isect.position;
        const float distSq = distance_squared(cps.isect.position, ts.ray.org);
        if (distSq <= 0.0f) { contribs.clear(); ts.stage = TS_DONE; return false; }
    }
    if (camDepth == ph.nCam - 1) {
        if (ph.lgtDepth == 1) {
            const float directLightPickProb = pick_light_prob(sc, sv.dlLight);
            perturb(sv.dlRndParam.x, offset, ts.offsetId);
            perturb(sv.dlRndParam.y, offset, ts.offsetId);
            direct_lighting(sc, camDepth, cps, ph.screenPos, directLightPickProb, sv, contribs);
        } else {
            const SurfaceVertex lv = lgtVerts[ph.nLgt - 1];
            connect_vertex(sc, camDepth, ph.nLgt - 1, ts.lps, lv, cps, sv, ph.screenPos, contribs);
        }
        ts.stage = TS_DONE; return false;
    }
    perturb(sv.bsdfRndParam.x, offset, ts.offsetId);
    perturb(sv.bsdfRndParam.y, offset, ts.offsetId);
    V3 bsdfContrib;
    if (!bsdf_sampling<false, true>(sc, cps, sv, cps, ts.ray.dir, bsdfContrib)) { ts.stage = TS_DONE; return false; }
    cps.throughput *= sv.rrWeight;
    ts.ray.org = cps.isect.position;
    ts.minT = LMC_ISECT_EPS; ts.maxT = dm_inf();
    ts.depth = camDepth + 1;
    return true;
}

// ---- GeneratePathBidir(scene, (-1,-1), minDepth, maxDepth, ...) ---------------------------------
// `ls` = the light subpath states (GenWork::ls); `sv` = the slot path.lgt[ph.nLgt] / path.cam[ph.nCam]
// AT ENTRY (the stage zero-fills it first, as the reference's emplace_back does).
LMC_HD bool gen_camera_begin(const Scene &sc, PathHead &ph, TraceState &ts, Rng &rng) {
    ph.screenPos.x = rng_uniform(rng); ph.screenPos.y = rng_uniform(rng);
    emit_from_camera(sc, ph.screenPos, ts.ray, ts.minT, ts.maxT, ts.cps);
    ts.stage = TS_G_CAM; ts.depth = 0;
    return true;
}

LMC_HD bool gen_stage_begin(const Scene &sc, PathHead &ph, TraceState &ts, BidirPathState *ls, Rng &rng) {
    ph.time = rng_uniform(rng);
    ts.nLightStates = 1;
    float lightPickProb = 1.0f;
    ph.lgtRndPos.x = rng_uniform(rng); ph.lgtRndPos.y = rng_uniform(rng);
    ph.lgtRndDir.x = rng_uniform(rng); ph.lgtRndDir.y = rng_uniform(rng);
    ph.lgtLight = pick_light(sc, rng_uniform(rng), lightPickProb);
    ph.lgtPrim = light_sample_discrete(sc, ph.lgtLight, rng_uniform(rng));
    emit_from_light(sc, lightPickProb, ph, ts.ray, ls[0]);
    ts.minT = LMC_ISECT_EPS; ts.maxT = dm_inf();
    ts.stage = TS_G_LGT; ts.depth = 0;
    return true;
}

template <class CL>
LMC_HD bool gen_stage_light(const Scene &sc, int minDepth, int maxDepth, PathHead &ph, SurfaceVertex &sv, TraceState &ts,
                            BidirPathState *ls, CL &contribs, Rng &rng, const Hit &hit) {
    const int lgtDepth = ts.depth;
    sv = surface_vertex_zero();
    ph.nLgt++;
    BidirPathState &cur = ls[lgtDepth];
    if (hit.tid < 0) { ts.nLightStates--; ph.nLgt--; return gen_camera_begin(sc, ph, ts, rng); }
    sv.tid = hit.tid;
    fill_isect(sc, ts.ray, hit, cur.isect, sv.st);
    sv.bsdfDiscrete = rng_uniform(rng);
    cur.wi = -ts.ray.dir;
    convert_mis(sc, lgtDepth, ph.lgtLight, ts.ray, cur);
    if (lgtDepth + 2 >= minDepth) connect_to_camera(sc, lgtDepth, cur, sv, ts.ray.org, contribs);
    if (maxDepth != -1 && lgtDepth + 2 >= maxDepth) return gen_camera_begin(sc, ph, ts, rng);
    ts.nLightStates++;
    sv.bsdfRndParam.x = rng_uniform(rng); sv.bsdfRndParam.y = rng_uniform(rng);
    V3 bsdfContrib;
    ls[lgtDepth + 1].ssJacobian = 0.0f;
    if (!bsdf_sampling<true, false>(sc, cur, sv, ls[lgtDepth + 1], ts.ray.dir, bsdfContrib)) {
        ts.nLightStates--; return gen_camera_begin(sc, ph, ts, rng);
    }
    if (!russian_roulette(lgtDepth, bsdfContrib, sv.rrWeight, ls[lgtDepth + 1].throughput, rng)) {
        ts.nLightStates--; return gen_camera_begin(sc, ph, ts, rng);
    }
    ts.ray.org = cur.isect.position;
    ts.depth = lgtDepth + 1;
    return true;
}

template <class CL>
LMC_HD bool gen_stage_camera(const Scene &sc, int minDepth, int maxDepth, PathHead &ph, SurfaceVertex &sv,
                             const SurfaceVertex *lgtVerts, TraceState &ts, const BidirPathState *ls, CL &contribs,
                             Rng &rng, const Hit &hit) {
    const int camDepth = ts.depth;
    BidirPathState &cps = ts.cps;
    sv = surface_vertex_zero();
    ph.nCam++;
    const bool hitSurface = hit.tid >= 0;
    if (hitSurface) { sv.tid = hit.tid; fill_isect(sc, ts.ray, hit, cps.isect, sv.st); }
    cps.wi = -ts.ray.dir;
    if (hitSurface) convert_mis(sc, camDepth, -1, ts.ray, cps);
    if (camDepth + 1 >= minDepth) {
        const int light = get_hit_light(sc, hitSurface, sv.tid);
        if (light >= 0) {
            handle_hit_light(sc, camDepth, light, hitSurface, ts.ray, ph.screenPos, cps, ph, contribs);
            ts.stage = TS_DONE; return false;
        }
    }
    if (!hitSurface || (maxDepth != -1 && camDepth + 1 >= maxDepth)) { ts.stage = TS_DONE; return false; }
    if (camDepth == 1) {
        ph.lensVertexPos = cps.isect.position;
        const float distSq = distance_squared(cps.isect.position, ts.ray.org);
        if (distSq <= 0.0f) { contribs.clear(); ts.stage = TS_DONE; return false; }
    }
    sv.bsdfDiscrete = rng_uniform(rng);
    if (camDepth + 2 >= minDepth) {
        float directLightPickProb = 1.0f;
        sv.dlLight = pick_light(sc, rng_uniform(rng), directLightPickProb);
        sv.dlRndParam.x = rng_uniform(rng); sv.dlRndParam.y = rng_uniform(rng);
        sv.dlPrim = light_sample_discrete(sc, sv.dlLight, rng_uniform(rng));
        direct_lighting(sc, camDepth, cps, ph.screenPos, directLightPickProb, sv, contribs);
    }
    int maxLgtDepth = ts.nLightStates - 1;
    if (maxDepth != -1) {
        const int m = maxDepth - camDepth - 3;
        if (m < maxLgtDepth) maxLgtDepth = m;
    }
    for (int lgtDepth = 0; lgtDepth <= maxLgtDepth; lgtDepth++) {
        if (camDepth + lgtDepth + 3 >= minDepth) {
            contribs.connect(sc, camDepth, lgtDepth, ls, lgtVerts, cps, sv, ph.screenPos);
        }
    }
    sv.bsdfRndParam.x = rng_uniform(rng); sv.bsdfRndParam.y = rng_uniform(rng);
    V3 bsdfContrib;
    if (!bsdf_sampling<false, false>(sc, cps, sv, cps, ts.ray.dir, bsdfContrib)) { ts.stage = TS_DONE; return false; }
    if (!russian_roulette(camDepth, bsdfContrib, sv.rrWeight, cps.throughput, rng)) { ts.stage = TS_DONE; return false; }
    ts.ray.org = cps.isect.position;
    ts.minT = LMC_ISECT_EPS; ts.maxT = dm_inf();
    ts.depth = camDepth + 1;
    return true;
}

}  // namespace lmc
