"""Per-vertex stages of the path function (tools/adgen/pathfn.py) as separately recorded programs.

The product does not generate one function per (camDepth, lightDepth) class like the reference
(42 classes x ~20 k statements).  The reverse sweep of a whole path is cut at the vertex boundaries: what
crosses a boundary is a small state (ray origin / direction, the two MIS accumulators, the throughput; at the
end of the light subpath also position / normal / wi), a stage's reverse function recomputes the stage from
its input state and turns the adjoints of its outputs into adjoints of its inputs.  Cutting there is exact for
the reference's sweep order as well: an adjoint overwrite at an `if` merge only discards what LATER statements
have accumulated, and those arrive as the stage's incoming output adjoints (tests compose the stages and
compare with the whole-path program).

Buffer conventions of a stage (the reference's serialized layout, SURVEY.md App. A.4, addressed relative to the
stage's first float):
  vert   the stage's own record: shape(46) [bsdfDiscrete, useAbsoluteParam, bsdf(10), rrWeight] | [light(56) ...]
  lvert  record of the LAST light-subpath vertex: shape(46), bsdfDiscrete, useAbsoluteParam, bsdf(10)
  scene  the 38-float scene block;  lightType  type of the emitting light (first light vertex only)
"""
import chadlike as cl
import pathfn as pf

NS = 11      # ray.org(3) ray.dir(3) accMISWPrev accMISWThis throughput(3)
NL = 14      # position(3) shadingNormal(3) wi(3) accMISWPrev accMISWThis throughput(3)

# name -> (number of differentiable inputs, number of outputs)
STAGES = {
    "emit_light": (4, NS),
    "light_vertex_first": (NS + 2, NS),
    "light_vertex": (NS + 2, NS),
    "light_last_first": (NS, NL),
    "light_last": (NS, NL),
    "connect_camera": (NL, 1),
    "emit_camera": (2, NS),
    "cam_vertex": (NS + 2, NS),
    "cam_hit_light": (NS, 1),
    "cam_hit_env": (NS, 1),
    "cam_direct": (NS + 2, 1),
    "cam_connect": (NS + NL, 1),
}


def _state_in(x):
    ray = pf.Ray()
    ray.org, ray.dir = x[0:3], x[3:6]
    ps = pf.PathState()
    ps.accMISWPrev, ps.accMISWThis, ps.throughput = x[6], x[7], x[8:11]
    return ray, ps


def _lps_in(x, lvert):
    ps = pf.PathState()
    ps.position, ps.shadingNormal, ps.wi = x[0:3], x[3:6], x[6:9]
    ps.accMISWPrev, ps.accMISWThis, ps.throughput = x[9], x[10], x[11:14]
    b = lvert + 2
    ps.geomNormal = pf.normalize(pf.cross(b.vec3(3), b.vec3(6)))
    return ps


def _state_out(ray, ps):
    return ray.org + ray.dir + [ps.accMISWPrev, ps.accMISWThis] + ps.throughput


def record_stage(name):
    nin, nout = STAGES[name]
    cl.begin_function()
    params = pf.Params()
    if name == "connect_camera":
        x = [cl.inp("lp[%d]" % i, True) for i in range(NL)]
    elif name == "cam_connect":
        x = [cl.inp("in[%d]" % i, True) for i in range(NS)] + [cl.inp("lp[%d]" % i, True) for i in range(NL)]
    else:
        x = [cl.inp("in[%d]" % i, True) for i in range(nin)]
    vert = pf.Buf(params, "vert")
    lvert = pf.Buf(params, "lvert")
    scn = pf.Scene(pf.Buf(params, "scene"))
    lightType = params.get("lightType", 0)
    if name == "emit_light":
        ray, ps = pf.Ray(), pf.PathState()
        pf.emit_from_light(vert + 1, scn, vert[0], x[0], x[1], x[2], x[3], ray, ps)
        out = _state_out(ray, ps)
    elif name in ("light_vertex_first", "light_vertex", "light_last_first", "light_last"):
        ray, ps = _state_in(x)
        buffer = pf.intersect(vert, ray, ps)
        bsdfDiscrete, useAbsoluteParam = buffer[0], buffer[1]
        buffer = buffer + 2
        ps.wi = pf.vneg(ray.dir)
        if name.endswith("_first"):
            pf.convert_mis_light_emit(lightType, ray, ps)
        else:
            pf.convert_mis(ray, ps)
        if name.startswith("light_last"):
            out = ps.position + ps.shadingNormal + ps.wi + [ps.accMISWPrev, ps.accMISWThis] + ps.throughput
        else:
            buffer, ray.dir = pf.bsdf_sampling(True, buffer, x[NS], x[NS + 1], bsdfDiscrete, useAbsoluteParam, ps)
            ps.throughput = pf.vmuls(ps.throughput, buffer[0])
            ray.org = ps.position
            out = _state_out(ray, ps)
    elif name == "connect_camera":
        ps = _lps_in(x, lvert)
        pf.connect_to_camera(scn, lvert + (pf.SER_SHAPE + 2), ps)
        out = [cl.log(pf.luminance(ps.throughput))]
    elif name == "emit_camera":
        ray, ps = pf.Ray(), pf.PathState()
        pf.emit_from_camera(scn, x[0], x[1], ray, ps)
        out = _state_out(ray, ps)
    elif name == "cam_vertex":
        ray, ps = _state_in(x)
        buffer = pf.intersect(vert, ray, ps)
        ps.wi = pf.vneg(ray.dir)
        pf.convert_mis(ray, ps)
        bsdfDiscrete, useAbsoluteParam = buffer[0], buffer[1]
        buffer = buffer + 2
        buffer, ray.dir = pf.bsdf_sampling(False, buffer, x[NS], x[NS + 1], bsdfDiscrete, useAbsoluteParam, ps)
        ps.throughput = pf.vmuls(ps.throughput, buffer[0])
        ray.org = ps.position
        out = _state_out(ray, ps)
    elif name == "cam_hit_light":
        ray, ps = _state_in(x)
        buffer = pf.intersect(vert, ray, ps)
        ps.wi = pf.vneg(ray.dir)
        pf.convert_mis_light_hit(buffer[0], ray, ps)
        pf.handle_hit_light(scn, buffer, ray.dir, ps)
        out = [cl.log(pf.luminance(ps.throughput))]
    elif name == "cam_hit_env":
        # The path left the scene and hit the environment light: there is no last surface vertex.  The reference still
        # differentiates an Intersect of whatever its reused buffer holds at that offset (src/path.cpp:2545-2548 skips the
        # shape block when shapeInst.obj == nullptr); nothing the env-light branches read depends on it, so its adjoint
        # is exactly zero whenever the stale values are finite.  This stage is that program without the dead Intersect.
        ray, ps = _state_in(x)
        ps.position = [pf.C(0.0)] * 3
        ps.geomNormal = [pf.C(0.0)] * 3
        ps.shadingNormal = [pf.C(0.0)] * 3
        buffer = vert + pf.SER_SHAPE
        ps.wi = pf.vneg(ray.dir)
        pf.convert_mis_light_hit(buffer[0], ray, ps)
        pf.handle_hit_light(scn, buffer, ray.dir, ps)
        out = [cl.log(pf.luminance(ps.throughput))]
    elif name == "cam_direct":
        ray, ps = _state_in(x)
        buffer = pf.intersect(vert, ray, ps)
        ps.wi = pf.vneg(ray.dir)
        pf.convert_mis(ray, ps)
        pf.direct_lighting(scn, buffer, ps, x[NS], x[NS + 1])
        out = [cl.log(pf.luminance(ps.throughput))]
    elif name == "cam_connect":
        ray, ps = _state_in(x)
        lps = _lps_in(x[NS:], lvert)
        buffer = pf.intersect(vert, ray, ps)
        ps.wi = pf.vneg(ray.dir)
        pf.convert_mis(ray, ps)
        pf.connect_vertex(buffer, lvert + (pf.SER_SHAPE + 2), lps, ps)
        out = [cl.log(pf.luminance(ps.throughput))]
    else:
        raise KeyError(name)
    f = cl.end_function()
    inputs = dict(params.nodes)
    for n in x:
        inputs[n.name] = n
    out = [cl.lift(o) for o in out]
    return cl.Program(f, inputs, out)


# ------------------------------------------------------------------------------------------------
# Python model of the C++ driver (csrc/core/pathgrad_rev.h): forward pass with checkpoints, reverse pass.
# ------------------------------------------------------------------------------------------------
_programs = {}


def program(name):
    if name not in _programs:
        _programs[name] = record_stage(name)
    return _programs[name]


def _run(name, scene, vert, voff, lvert_off, lightType, xin, out_adj=None, compat=True):
    prog = program(name)
    vals = {}
    lpbase = NS if name == "cam_connect" else 0
    for key in prog.inputs:
        if key.startswith("in["):
            vals[key] = xin[int(key[3:-1])]
        elif key.startswith("lp["):
            vals[key] = xin[lpbase + int(key[3:-1])]
        elif key.startswith("vert["):
            vals[key] = float(vert[voff + int(key[5:-1])])
        elif key.startswith("lvert["):
            vals[key] = float(vert[lvert_off + int(key[6:-1])])
        elif key.startswith("scene["):
            vals[key] = float(scene[int(key[6:-1])])
        elif key.startswith("lightType"):
            vals[key] = float(lightType)
    outs, adj = prog.evaluate(vals, out_adj, compat)
    if adj is not None:
        nin = STAGES[name][0]
        if name == "connect_camera":
            adj = [adj.get("lp[%d]" % i, 0.0) for i in range(NL)]
        elif name == "cam_connect":
            adj = [adj.get("in[%d]" % i, 0.0) for i in range(NS)] + [adj.get("lp[%d]" % i, 0.0) for i in range(NL)]
        else:
            adj = [adj.get("in[%d]" % i, 0.0) for i in range(nin)]
    return outs, adj


def plan(c, l, env_hit=False):
    """[(stage name, offset of its record in vertParams, number of PSS values it consumes)] + offset of lvert."""
    off = 3
    steps = []
    lvert = -1
    if l > 1:
        steps.append(("emit_light", off, 4))
        off += 1 + pf.SER_LIGHT
        for d in range(l - 1):
            first = "_first" if d == 0 else ""
            if d == l - 2:
                steps.append(("light_last" + first, off, 0))
                lvert = off
                off += pf.SER_SHAPE + 2 + pf.SER_BSDF
                if c == 1:
                    steps.append(("connect_camera", lvert, 0))
            else:
                steps.append(("light_vertex" + first, off, 2))
                off += pf.SER_SHAPE + 2 + pf.SER_BSDF + 1
    if c > 1:
        steps.append(("emit_camera", off, 2))
        for d in range(c - 1):
            if d == c - 2:
                if l == 0:
                    steps.append(("cam_hit_env" if env_hit else "cam_hit_light", off, 0))
                elif l == 1:
                    steps.append(("cam_direct", off, 2))
                else:
                    steps.append(("cam_connect", off, 0))
            else:
                steps.append(("cam_vertex", off, 2))
                off += pf.SER_SHAPE + 2 + pf.SER_BSDF + 1
    return steps, lvert


def path_grad_staged(c, l, scene, primary, vert, compat=True):
    env_hit = False
    if l == 0 and c > 1:
        env_hit = vert[3 + (c - 2) * (pf.SER_SHAPE + 2 + pf.SER_BSDF + 1) + pf.SER_SHAPE] == pf.LIGHT_ENV
    steps, lvert = plan(c, l, env_hit)
    dim = 2 * max(c + l - 1, 2)
    pss = [float(v) for v in primary[1:dim + 1]]
    lightType = vert[3 + 1] if l > 1 else 0.0
    # forward with checkpoints
    ck = []
    state, lps = None, None
    pi = 0
    value = None
    for (name, off, npss) in steps:
        if name == "emit_light" or name == "emit_camera":
            xin = pss[pi:pi + npss]
        elif name == "connect_camera":
            xin = lps
        elif name == "cam_connect":
            xin = state + lps
        else:
            xin = state + pss[pi:pi + npss]
        ck.append((name, off, pi, npss, xin))
        pi += npss
        outs, _ = _run(name, scene, vert, off, lvert, lightType, xin)
        if name.startswith("light_last"):
            lps = outs
        elif STAGES[name][1] == 1:
            value = outs[0]
        else:
            state = outs
    # reverse
    grad = [0.0] * dim
    adj_state, adj_lps = None, None
    for (name, off, p0, npss, xin) in reversed(ck):
        nout = STAGES[name][1]
        if nout == 1:
            out_adj = [1.0]
        elif name.startswith("light_last"):
            out_adj = adj_lps
        else:
            out_adj = adj_state
        _, adj = _run(name, scene, vert, off, lvert, lightType, xin, out_adj, compat)
        if name in ("emit_light", "emit_camera"):
            for k in range(npss):
                grad[p0 + k] = adj[k]
            adj_state = None
        elif name == "connect_camera":
            adj_lps = adj
        elif name == "cam_connect":
            adj_state, adj_lps = adj[:NS], adj[NS:]
        else:
            adj_state = adj[:NS]
            for k in range(npss):
                grad[p0 + k] = adj[NS + k]
    return value, grad
