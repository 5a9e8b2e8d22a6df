"""The path-contribution function of a (camDepth, lightDepth) class, recorded with tools/adgen/chadlike.py.

Statement structure follows the reference's AD twin of the path sampler (Static mode) so that node
identity and the position of every `if` merge -- which is what the reference's reverse sweep is
sensitive to, SURVEY.md App. B#13 -- are the reference's:
  RegisterPathFuncBidirMALA                      src/path.cpp:3664-3911
  EmitFromLight / ConvertMIS* / ConnectToCamera  src/path.cpp:2799-2958
  BSDFSampling / EmitFromCamera / HandleHitLight / DirectLighting / ConnectVertex   src/path.cpp:2960-3380
  EvaluateBSDF / SampleBSDF dispatch             src/bsdf.cpp:13-171
  Lambertian / Phong / RoughDielectric twins     src/lambertian.cpp:95-151, src/phong.cpp:171-393, src/roughdielectric.cpp:332-528
  microfacet terms                               src/microfacet.h
  light dispatch and twins                       src/light.cpp, src/envlight.cpp:250-399, src/arealight.cpp:106-208, src/pointlight.cpp:74-116
  shapes                                         src/shape.cpp:13-122, src/trianglemesh.cpp:81-105,313-327,367-473
  camera                                         src/camera.cpp:53-66
The lens-contribution statements of the twin are not recorded: in Static mode nothing that reaches the
dependent variable reads them, and a merge whose outputs carry no adjoint emits nothing (src/chad.cpp:243-252).
Buffers are addressed exactly like the reference's serialized layouts.
"""
import math

from chadlike import (And, Eq, Gt, Gte, Lt, Lte, acos, atan2, begin_else, begin_else_if, begin_if, const, cos, dot3,
                      end_if, exp, fabs, fmax, if_else, inp, inverse, length3, log, pow_, set_cond_output, sin, sqrt,
                      square)

PI = math.pi
INVPI = 1.0 / math.pi
TWOPI = 2.0 * math.pi
INVTWOPI = 1.0 / TWOPI
INVFOURPI = 1.0 / (4.0 * math.pi)
PIOVERTWO = 0.5 * math.pi
PIOVERFOUR = 0.25 * math.pi
AD_EPS = 1e-6

SER_SHAPE, SER_BSDF, SER_LIGHT, SER_SCENE = 46, 10, 56, 38
BSDF_LAMBERTIAN, BSDF_PHONG, BSDF_ROUGHDIELECTRIC = 0.0, 1.0, 2.0
LIGHT_POINT, LIGHT_AREA, LIGHT_ENV = 0.0, 1.0, 2.0
SHAPE_TRIANGLEMESH = 0.0


class Params:
    """Named non-differentiable inputs (the serialized buffers)."""

    def __init__(self):
        self.nodes = {}

    def get(self, name, i):
        k = "%s[%d]" % (name, i)
        if k not in self.nodes:
            self.nodes[k] = inp(k, False)
        return self.nodes[k]


class Buf:
    def __init__(self, params, name, off=0):
        self.params, self.name, self.off = params, name, off

    def __getitem__(self, i):
        return self.params.get(self.name, self.off + i)

    def __add__(self, k):
        return Buf(self.params, self.name, self.off + k)

    def vec3(self, i):
        return [self[i], self[i + 1], self[i + 2]]


# ---- small vectors (lists of nodes) ----------------------------------------------------------
def C(v): return const(v)
def vadd(a, b): return [a[i] + b[i] for i in range(3)]
def vsub(a, b): return [a[i] - b[i] for i in range(3)]
def vneg(a): return [-a[i] for i in range(3)]
def vmuls(a, s): return [a[i] * s for i in range(3)]           # v * s
def smulv(s, a): return [s * a[i] for i in range(3)]           # s * v
def cwise(a, b): return [a[i] * b[i] for i in range(3)]
def dot(a, b): return dot3(a, b)
def length_squared(v): return square(v[0]) + square(v[1]) + square(v[2])
def distance_squared(a, b): return square(a[0] - b[0]) + square(a[1] - b[1]) + square(a[2] - b[2])
def normalize(v): return vmuls(v, inverse(length3(v)))
def luminance(v): return v[0] * 0.212671 + v[1] * 0.715160 + v[2] * 0.072169
def MIS(x): return square(x)


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def reflect(wi, n):
    return vsub(smulv(2.0 * dot(wi, n), n), wi)


def refract(wi, n, cosThetaT, eta, invEta):
    eta_ = if_else(Lt(cosThetaT, 0.0), invEta, eta)
    return vsub(vmuls(n, dot(wi, n) * eta_ + cosThetaT), vmuls(wi, eta_))


def coordinate_system(n):                # src/utils.h:234-262
    ret = begin_if(Lt(n[2], -1.0 + 1e-6), 6)
    set_cond_output([C(0.0), C(-1.0), C(0.0), C(-1.0), C(0.0), C(0.0)])
    begin_else()
    a = 1.0 / (1.0 + n[2])
    b = -n[0] * n[1] * a
    b1 = [1.0 - square(n[0]) * a, b, -n[0]]
    b2 = [b, 1.0 - square(n[1]) * a, -n[1]]
    set_cond_output(b1 + b2)
    end_if()
    return ret[0:3], ret[3:6]


def tent(s):                             # src/utils.h:269-277
    ret = begin_if(Lt(s, 0.5), 1)
    set_cond_output([1.0 - sqrt(2.0 * s)])
    begin_else()
    set_cond_output([sqrt(2.0 * (s - 0.5)) - 1.0])
    end_if()
    return ret[0]


# ---- transforms (parameters only) --------------------------------------------------------------
def static_matrix(b):
    """AnimatedTransform (15 floats) -> Translate(translate[0]) * ToMatrix4x4(rotate[0]), row-major 4x4
    (src/animatedtransform.cpp:65-67, src/quaternion.h:13-38: note the transpose)."""
    t = [b[1], b[2], b[3]]
    q = [b[7], b[8], b[9], b[10]]
    xx, yy, zz = q[0] * q[0], q[1] * q[1], q[2] * q[2]
    xy, xz, yz = q[0] * q[1], q[0] * q[2], q[1] * q[2]
    wx, wy, wz = q[0] * q[3], q[1] * q[3], q[2] * q[3]
    m = [[None] * 4 for _ in range(4)]
    m[0][0] = 1.0 - 2.0 * (yy + zz); m[1][0] = 2.0 * (xy + wz); m[2][0] = 2.0 * (xz - wy); m[3][0] = C(0.0)
    m[0][1] = 2.0 * (xy - wz); m[1][1] = 1.0 - 2.0 * (xx + zz); m[2][1] = 2.0 * (yz + wx); m[3][1] = C(0.0)
    m[0][2] = 2.0 * (xz + wy); m[1][2] = 2.0 * (yz - wx); m[2][2] = 1.0 - 2.0 * (xx + yy); m[3][2] = C(0.0)
    m[0][3] = t[0]; m[1][3] = t[1]; m[2][3] = t[2]; m[3][3] = C(1.0)
    return m


def xform_vector(m, v):
    return [m[r][0] * v[0] + m[r][1] * v[1] + m[r][2] * v[2] for r in range(3)]


def xform_point(m, p):
    x, y, z, w = [m[r][0] * p[0] + m[r][1] * p[1] + m[r][2] * p[2] + m[r][3] for r in range(4)]
    iw = inverse(w)
    return [x * iw, y * iw, z * iw]


class Scene:
    def __init__(self, s):
        self.useLightCoordinateSampling = s[0]
        self.sampleToCam = [[s[1 + c * 4 + r] for c in range(4)] for r in range(4)]   # column-major storage
        self.camToWorld = static_matrix(s + 17)
        self.screenPixelCount = s[32]
        self.camDist = s[33]
        self.bsphereCenter = s.vec3(34)
        self.bsphereRadius = s[37]


class Ray:
    def __init__(self):
        self.org = self.dir = None


class PathState:
    def __init__(self):
        self.position = self.geomNormal = self.shadingNormal = None
        self.wi = None
        self.accMISWPrev = self.accMISWThis = None
        self.throughput = None


def sample_primary(scn, sx, sy):         # src/camera.cpp:53-66 (static)
    o = xform_point(scn.sampleToCam, [sx, sy, C(0.0)])
    d = normalize(o)
    org = xform_point(scn.camToWorld, [C(0.0), C(0.0), C(0.0)])
    return org, xform_vector(scn.camToWorld, d)


# ---- shapes --------------------------------------------------------------------------------------
def triangle_intersect(ray, p0, e1, e2, n0, n1, n2):     # src/trianglemesh.cpp:81-105
    geomNormal = normalize(cross(e1, e2))
    s1 = cross(ray.dir, e2)
    divisor = dot(s1, e1)
    invDivisor = inverse(divisor)
    s = vsub(ray.org, p0)
    u = dot(s, s1) * invDivisor
    s2 = cross(s, e1)
    v = dot(ray.dir, s2) * invDivisor
    t = dot(e2, s2) * invDivisor
    w = 1.0 - u - v
    position = vadd(ray.org, smulv(t, ray.dir))
    shadingNormal = normalize(vadd(vadd(smulv(w, n0), smulv(u, n1)), smulv(v, n2)))
    return position, geomNormal, shadingNormal


def intersect(buffer, ray, ps):          # src/shape.cpp:13-66 + IntersectTriangleMesh (static)
    ret = begin_if(Eq(buffer[0], SHAPE_TRIANGLEMESH), 9)
    b = buffer + 2
    pos, gn, sn = triangle_intersect(ray, b.vec3(0), b.vec3(3), b.vec3(6), b.vec3(9), b.vec3(12), b.vec3(15))
    set_cond_output(pos + gn + sn)
    begin_else()
    set_cond_output([C(0.0)] * 9)
    end_if()
    ps.position, ps.geomNormal, ps.shadingNormal = ret[0:3], ret[3:6], ret[6:9]
    return buffer + SER_SHAPE


def sample_shape(buffer, r0, r1):        # src/shape.cpp:68-105, src/trianglemesh.cpp:313-327
    ret = begin_if(Eq(buffer[0], SHAPE_TRIANGLEMESH), 7)
    b = buffer + 2
    p0, e1, e2, n0, n1, n2 = b.vec3(0), b.vec3(3), b.vec3(6), b.vec3(9), b.vec3(12), b.vec3(15)
    a = sqrt((1.0 + AD_EPS) - r0)
    b1 = 1.0 - a
    b2 = a * r1
    pos = vadd(vadd(p0, vmuls(e1, b1)), vmuls(e2, b2))
    nrm = normalize(vadd(vadd(vmuls(n0, 1.0 - b1 - b2), vmuls(n1, b1)), vmuls(n2, b2)))
    set_cond_output(pos + nrm + [buffer[SER_SHAPE - 1]])
    begin_else()
    set_cond_output([C(0.0)] * 7)
    end_if()
    return ret[0:3], ret[3:6], ret[6]


def sample_shape_pdf(buffer):            # src/shape.cpp:107-122
    ret = begin_if(Eq(buffer[0], SHAPE_TRIANGLEMESH), 1)
    set_cond_output([buffer[SER_SHAPE - 1]])
    begin_else()
    set_cond_output([C(0.0)])
    end_if()
    return ret[0]


# ---- microfacet ----------------------------------------------------------------------------------
def beckmann_D(localH, alphaU, alphaV):
    cosTheta2 = square(localH[2])
    e = (square(localH[0]) / square(alphaU) + square(localH[1]) / square(alphaV)) / cosTheta2
    return exp(-e) / (PI * alphaU * alphaV * square(cosTheta2))


def beckmann_G1(alpha, cosTheta):        # src/microfacet.h BeckmennGeometryTerm(ADFloat)
    tanTheta = sqrt(fabs((1.0 + 1e-6) - square(cosTheta))) / cosTheta
    ret = begin_if(Lte(tanTheta, 0.0), 1)
    set_cond_output([C(1.0)])
    begin_else()
    a = inverse(alpha * tanTheta)
    g = begin_if(Gte(a, 1.6), 1)
    set_cond_output([C(1.0)])
    begin_else()
    aSqr = square(a)
    set_cond_output([(3.535 * a + 2.181 * aSqr) / (1.0 + 2.276 * a + 2.577 * aSqr)])
    end_if()
    set_cond_output([g[0]])
    end_if()
    return ret[0]


def beckmann_G(alpha, cosWi, cosWo):
    g0 = beckmann_G1(alpha, cosWi)
    g1 = beckmann_G1(alpha, cosWo)
    return g0 * g1


def fresnel_ext(cosThetaI_, eta, invEta, want_t):
    scale = if_else(Gt(cosThetaI_, 0.0), invEta, eta)
    cosThetaTSqr = 1.0 - (1.0 - square(cosThetaI_)) * square(scale)
    if want_t:
        ret = begin_if(Lte(cosThetaTSqr, 0.0), 2)
        set_cond_output([C(0.0), C(1.0)])
    else:
        ret = begin_if(Lte(cosThetaTSqr, 0.0), 1)
        set_cond_output([C(1.0)])
    begin_else()
    cosThetaI = fabs(cosThetaI_)
    cosThetaT = sqrt(cosThetaTSqr)
    etaCosThetaT = eta * cosThetaT
    etaCosThetaI = eta * cosThetaI
    Rs = (cosThetaI - etaCosThetaT) / (cosThetaI + etaCosThetaT)
    Rp = (etaCosThetaI - cosThetaT) / (etaCosThetaI + cosThetaT)
    if want_t:
        cosThetaT_ = if_else(Gt(cosThetaI_, 0.0), -cosThetaT, cosThetaT)
        F = 0.5 * (square(Rs) + square(Rp))
        set_cond_output([cosThetaT_, F])
    else:
        F = 0.5 * (square(Rs) + square(Rp))
        set_cond_output([F])
    end_if()
    if want_t:
        return ret[1], ret[0]
    return ret[0], None


def sample_micronormal(r0, r1, alpha):   # src/microfacet.h SampleMicronormal<ADFloat>
    phiM = TWOPI * r1
    sinPhiM = sin(phiM)
    cosPhiM = cos(phiM)
    alphaSqr = square(alpha)
    tanThetaMSqr = alphaSqr * (-log(fmax(1.0 - r0, 1e-6)))
    cosThetaM = 1.0 / sqrt(1.0 + tanThetaMSqr)
    cosThetaMSqr = square(cosThetaM)
    pdfW = (1.0 - r0) / (PI * alphaSqr * cosThetaM * cosThetaMSqr)
    sinThetaMSq = fmax(1.0 - cosThetaMSqr, AD_EPS)
    sinThetaM = sqrt(sinThetaMSq)
    return [sinThetaM * cosPhiM, sinThetaM * sinPhiM, cosThetaM], pdfW


def sample_cos_hemisphere(r0, r1):       # src/sampling.h:103-110
    phi = TWOPI * r0
    tmp = sqrt(fmax(1.0 - r1, AD_EPS))
    return [cos(phi) * tmp, sin(phi) * tmp, sqrt(fmax(r1, AD_EPS))]


def sample_sphere(c0, c1):               # src/sampling.h:6-16
    scaledTheta = TWOPI * c0
    scaledPhi = PI * c1
    sinPhi = sin(scaledPhi)
    cosPhi = cos(scaledPhi)
    d = [sinPhi * cos(scaledTheta), sinPhi * sin(scaledTheta), cosPhi]
    jacobian = fabs(sinPhi) * TWOPI * PI
    return d, jacobian


def sample_concentric_disc(r0, r1):      # src/sampling.h:72-101
    a1 = 2.0 * r0 - 1.0
    a2 = 2.0 * r1 - 1.0
    ret = begin_if(Eq(a1, 0.0), 2)
    set_cond_output([C(0.0), C(0.0)])
    begin_else_if(Eq(a2, 0.0))
    set_cond_output([C(0.0), C(0.0)])
    begin_else_if(Gt(square(a1), square(a2)))
    set_cond_output([a1, PIOVERFOUR * (a2 / a1)])
    begin_else()
    set_cond_output([a2, PIOVERTWO - (a1 / a2) * PIOVERFOUR])
    end_if()
    r, phi = ret
    sinPhi = sin(phi)
    cosPhi = cos(phi)
    return r * cosPhi, r * sinPhi


# ---- BSDFs ---------------------------------------------------------------------------------------
def face_normal(normal, cosWi):
    ret = begin_if(Gt(cosWi, 0.0), 4)
    set_cond_output([normal[0], normal[1], normal[2], cosWi])
    begin_else()
    set_cond_output([-normal[0], -normal[1], -normal[2], -cosWi])
    end_if()
    return ret[0:3], ret[3]


def evaluate_lambertian(b, wi, normal, wo):
    Kd = b.vec3(0)
    cosWi = dot(normal, wi)
    normal_, cosWi = face_normal(normal, cosWi)
    cosWo = dot(normal_, wo)
    fwdScalar = cosWo * INVPI
    revScalar = cosWi * INVPI
    contrib = smulv(fwdScalar, Kd)
    return contrib, cosWo, fwdScalar, revScalar


def sample_lambertian(b, wi, normal, r0, r1):
    Kd = b.vec3(0)
    cosWi = dot(wi, normal)
    normal_, cosWi = face_normal(normal, cosWi)
    # Sample(normal_, rndParam, wo, cosWo, pdf), src/lambertian.cpp
    b0, b1 = coordinate_system(normal_)
    ret = sample_cos_hemisphere(r0, r1)
    wo = vadd(vadd(smulv(ret[0], b0), smulv(ret[1], b1)), smulv(ret[2], normal_))
    cosWo = ret[2]
    pdf = ret[2] * INVPI
    revPdf = cosWi * INVPI
    return wo, Kd, cosWo, pdf, revPdf


def evaluate_phong(b, wi, normal, wo):
    Kd, Ks, exponent, KsWeight = b.vec3(0), b.vec3(3), b[6], b[7]
    cosWi = dot(normal, wi)
    normal_, cosWi = face_normal(normal, cosWi)
    cosWo = dot(normal_, wo)
    # `alpha` is created inside the KsWeight > 0 branch in the reference; it has no `if` of its own, so
    # recording it just before the branch is the same program for the sweep
    contrib, pdf, revPdf = _phong_lobes_lazy(Kd, Ks, exponent, KsWeight, lambda: dot(reflect(wi, normal_), wo), cosWo, cosWi)
    return contrib, cosWo, pdf, revPdf


def _phong_lobes_lazy(Kd, Ks, exponent, KsWeight, alpha_fn, cosWo, cosWi):
    ret = begin_if(Gt(KsWeight, 0.0), 4)
    alpha = alpha_fn()
    weight = pow_(alpha, exponent) * INVTWOPI
    r2 = begin_if(Gt(weight, 1e-10), 4)
    expoConst1 = exponent + 1.0
    expoConst2 = exponent + 2.0
    specContrib = vmuls(Ks, expoConst2 * weight)
    specPdf = KsWeight * expoConst1 * weight
    set_cond_output(specContrib + [specPdf])
    begin_else()
    set_cond_output([C(0.0)] * 4)
    end_if()
    set_cond_output(r2)
    begin_else()
    set_cond_output([C(0.0)] * 4)
    end_if()
    contrib = ret[0:3]
    pdf = ret[3]
    revPdf = ret[3]
    ret = begin_if(Lt(KsWeight, 1.0), 5)
    diffContrib = vmuls(Kd, C(INVPI))
    tmp = (1.0 - KsWeight) * INVPI
    diffPdf = tmp * cosWo
    revDiffPdf = tmp * cosWi
    set_cond_output(diffContrib + [diffPdf, revDiffPdf])
    begin_else()
    set_cond_output([C(0.0)] * 5)
    end_if()
    contrib = [contrib[i] + ret[i] for i in range(3)]
    pdf = pdf + ret[3]
    revPdf = revPdf + ret[4]
    contrib = vmuls(contrib, cosWo)
    return contrib, pdf, revPdf


def sample_phong(b, wi, normal, r0, r1, uDiscrete):
    Kd, Ks, exponent, KsWeight = b.vec3(0), b.vec3(3), b[6], b[7]
    cosWi = dot(normal, wi)
    normal_, cosWi = face_normal(normal, cosWi)
    R = reflect(wi, normal_)
    ret = begin_if(Gt(uDiscrete, KsWeight), 4)
    localDir = sample_cos_hemisphere(r0, r1)
    b0, b1 = coordinate_system(normal_)
    wo = vadd(vadd(smulv(localDir[0], b0), smulv(localDir[1], b1)), smulv(localDir[2], normal_))
    set_cond_output(wo + [1.0 - KsWeight])
    begin_else()
    power = 1.0 / (exponent + 1.0)
    cosAlpha = pow_(r1, power)
    sinAlpha = sqrt(fmax(1.0 - square(cosAlpha), 1e-6))
    phi = TWOPI * r0
    localDir = [sinAlpha * cos(phi), sinAlpha * sin(phi), cosAlpha]
    b0, b1 = coordinate_system(R)
    wo = vadd(vadd(smulv(localDir[0], b0), smulv(localDir[1], b1)), smulv(localDir[2], R))
    set_cond_output(wo + [KsWeight])
    end_if()
    wo = ret[0:3]
    cosWo = dot(normal_, wo)
    contrib, pdf, revPdf = _phong_lobes_lazy(Kd, Ks, exponent, KsWeight, lambda: dot(R, wo), cosWo, cosWi)
    contrib = vmuls(contrib, inverse(pdf))
    return wo, contrib, cosWo, pdf, revPdf


def evaluate_roughdielectric(adjoint, b, wi, normal, wo):
    Ks, Kt, eta, invEta, alpha = b.vec3(0), b.vec3(3), b[6], b[7], b[8]
    cosWi = dot(wi, normal)
    cosWo = dot(wo, normal)
    side = cosWi * cosWo
    eta_ = if_else(Gt(cosWi, 0.0), eta, invEta)
    revEta_ = if_else(Gt(cosWo, 0.0), eta, invEta)
    reflect_c = Gt(side, 0.0)
    vH = begin_if(reflect_c, 3)
    set_cond_output(normalize(vadd(wi, wo)))
    begin_else()
    set_cond_output(normalize(vadd(wi, vmuls(wo, eta_))))
    end_if()
    H = vH
    Hside = dot(H, normal)
    vH = begin_if(Lt(Hside, 0.0), 3)
    set_cond_output(vneg(H))
    begin_else()
    set_cond_output(H)
    end_if()
    H = vH
    cosHWi = dot(wi, H)
    cosHWo = dot(wo, H)
    b0, b1 = coordinate_system(normal)
    localH = [dot(b0, H), dot(b1, H), dot(normal, H)]
    D = beckmann_D(localH, alpha, alpha)
    revCosHWi = cosHWo
    revCosHWo = cosHWi
    F, _ = fresnel_ext(cosHWi, eta, invEta, False)
    aCosWi = fabs(cosWi)
    aCosWo = fabs(cosWo)
    Gterm = beckmann_G(alpha, aCosWi, aCosWo)
    scaledAlpha = alpha * (1.2 - 0.2 * sqrt(aCosWi))
    scaledD = beckmann_D(localH, scaledAlpha, scaledAlpha)
    prob = localH[2] * scaledD
    revScaledAlpha = alpha * (1.2 - 0.2 * sqrt(aCosWo))
    revScaledD = beckmann_D(localH, revScaledAlpha, revScaledAlpha)
    revProb = localH[2] * revScaledD
    ret = begin_if(reflect_c, 5)
    scalar = fabs(F * D * Gterm / (4.0 * cosWi))
    contrib = vmuls(Ks, scalar)
    pdf = fabs(prob * F / (4.0 * cosHWo))
    revPdf = fabs(revProb * F / (4.0 * revCosHWo))
    set_cond_output(contrib + [pdf, revPdf])
    begin_else()
    sqrtDenom = cosHWi + eta_ * cosHWo
    revSqrtDenom = revCosHWi + revEta_ * revCosHWo
    factor = C(1.0) if adjoint else square(inverse(eta_))
    scalar = fabs(factor * ((1.0 - F) * D * Gterm * square(eta_) * cosHWi * cosHWo) / (cosWi * square(sqrtDenom)))
    contrib = vmuls(Kt, scalar)
    pdf = fabs(prob * (1.0 - F) * (square(eta_) * cosHWo) / (square(sqrtDenom)))
    revPdf = fabs(revProb * (1.0 - F) * (square(revEta_) * revCosHWo) / (square(revSqrtDenom)))
    set_cond_output(contrib + [pdf, revPdf])
    end_if()
    return ret[0:3], cosWo, ret[3], ret[4]


def sample_roughdielectric(adjoint, b, wi, normal, r0, r1, uDiscrete):
    Ks, Kt, eta, invEta, alpha = b.vec3(0), b.vec3(3), b[6], b[7], b[8]
    cosWi = dot(wi, normal)
    scaledAlpha = alpha * (1.2 - 0.2 * sqrt(fabs(cosWi)))
    localH, mPdf = sample_micronormal(r0, r1, scaledAlpha)
    b0, b1 = coordinate_system(normal)
    H = vadd(vadd(smulv(localH[0], b0), smulv(localH[1], b1)), smulv(localH[2], normal))
    cosHWi = dot(wi, H)
    F, cosThetaT = fresnel_ext(cosHWi, eta, invEta, True)
    ret = begin_if(Lte(uDiscrete, F), 9)
    wo = reflect(wi, H)
    refl = Ks
    cosHWo = dot(wo, H)
    pdf = fabs(mPdf * F / (4.0 * cosHWo))
    revCosHWo = cosHWi
    rev_dwh_dwo = inverse(4.0 * revCosHWo)
    cosWo = dot(wo, normal)
    revScaledAlp = alpha * (1.2 - 0.2 * sqrt(fabs(cosWo)))
    revD = beckmann_D(localH, revScaledAlp, revScaledAlp)
    revPdf = fabs(F * revD * localH[2] * rev_dwh_dwo)
    set_cond_output(wo + refl + [cosWo, pdf, revPdf])
    begin_else()
    wo = refract(wi, H, cosThetaT, eta, invEta)
    eta_ = if_else(Gt(cosWi, 0.0), eta, invEta)
    factor = C(1.0) if adjoint else square(inverse(eta_))
    refl = vmuls(Kt, factor)
    cosHWo = dot(wo, H)
    sqrtDenom = cosHWi + eta_ * cosHWo
    dwh_dwo = (square(eta_) * cosHWo) / square(sqrtDenom)
    pdf = fabs(mPdf * (1.0 - F) * fabs(dwh_dwo))
    cosWo = dot(wo, normal)
    revEta_ = if_else(Gt(cosWo, 0.0), eta, invEta)
    revCosHWi = cosHWo
    revCosHWo = cosHWi
    revSqrtDenom = revCosHWi + revEta_ * revCosHWo
    rev_dwh_dwo = (square(revEta_) * revCosHWo) / square(revSqrtDenom)
    revScaledAlp = alpha * (1.2 - 0.2 * sqrt(fabs(cosWo)))
    revD = beckmann_D(localH, revScaledAlp, revScaledAlp)
    revPdf = fabs((1.0 - F) * revD * localH[2] * rev_dwh_dwo)
    set_cond_output(wo + refl + [cosWo, pdf, revPdf])
    end_if()
    wo, refl, cosWo, pdf, revPdf = ret[0:3], ret[3:6], ret[6], ret[7], ret[8]
    aCosWi = fabs(cosWi)
    aCosWo = fabs(cosWo)
    D = beckmann_D(localH, alpha, alpha)
    Gterm = beckmann_G(alpha, aCosWi, aCosWo)
    numerator = D * Gterm * cosHWi
    denominator = mPdf * aCosWi
    contrib = vmuls(refl, fabs(numerator / denominator))
    return wo, contrib, cosWo, pdf, revPdf


# The BSDF table of the differentiable twin, in the order of the reference's if-chain (src/bsdf.cpp:24-55, 87-150):
# (type id, evaluate(adjoint, b, wi, normal, wo) -> contrib, cosWo, pdf, revPdf,
#           sample(adjoint, b, wi, normal, r0, r1, uDiscrete) -> wo, contrib, cosWo, pdf, revPdf).
# A new BSDF is a new row here (its adjoint is generated) plus a row of LMC_BSDF_TABLE in csrc/core/bsdf.h.
BSDF_TABLE = [
    (BSDF_PHONG, lambda adj, b, wi, n, wo: evaluate_phong(b, wi, n, wo),
     lambda adj, b, wi, n, r0, r1, u: sample_phong(b, wi, n, r0, r1, u)),
    (BSDF_ROUGHDIELECTRIC, lambda adj, b, wi, n, wo: evaluate_roughdielectric(adj, b, wi, n, wo),
     lambda adj, b, wi, n, r0, r1, u: sample_roughdielectric(adj, b, wi, n, r0, r1, u)),
    (BSDF_LAMBERTIAN, lambda adj, b, wi, n, wo: evaluate_lambertian(b, wi, n, wo),
     lambda adj, b, wi, n, r0, r1, u: sample_lambertian(b, wi, n, r0, r1)),
]


def evaluate_bsdf(adjoint, buffer, wi, normal, wo):     # src/bsdf.cpp:13-66
    t = buffer[0]
    b = buffer + 1
    ret = None
    for k, (tid, ev, _) in enumerate(BSDF_TABLE):
        if k == 0:
            ret = begin_if(Eq(t, tid), 6)
        else:
            begin_else_if(Eq(t, tid))
        c, cw, p, rp = ev(adjoint, b, wi, normal, wo)
        set_cond_output(c + [cw, p, rp])
    begin_else()
    set_cond_output([C(0.0)] * 6)
    end_if()
    return ret[0:3], ret[3], ret[4], ret[5]


def sample_bsdf(adjoint, buffer, wi, normal, r0, r1, uDiscrete):     # src/bsdf.cpp:68-171
    t = buffer[0]
    b = buffer + 1
    ret = None
    for k, (tid, _, sm) in enumerate(BSDF_TABLE):
        if k == 0:
            ret = begin_if(Eq(t, tid), 9)
        else:
            begin_else_if(Eq(t, tid))
        wo, c, cw, p, rp = sm(adjoint, b, wi, normal, r0, r1, uDiscrete)
        set_cond_output(wo + c + [cw, p, rp])
    begin_else()
    set_cond_output([C(0.0)] * 9)
    end_if()
    return ret[0:3], ret[3:6], ret[6], ret[7], ret[8]


def shading_normal_correction_adjoint(wi, ps, wo):      # src/path.cpp:56-70
    cosWi = dot(ps.shadingNormal, wi)
    cosWo = dot(ps.shadingNormal, wo)
    wiDotGeoN = dot(ps.geomNormal, wi)
    woDotGeoN = dot(ps.geomNormal, wo)
    return fabs((woDotGeoN * cosWi) / (wiDotGeoN * cosWo))


# ---- lights --------------------------------------------------------------------------------------
class EnvRec:
    def __init__(self, buffer):
        b = buffer + 1
        self.toWorld = static_matrix(b)
        self.toLight = static_matrix(b + 15)
        b = b + 30
        self.cdfCol0, self.cdfCol1, self.cdfRow0, self.cdfRow1, self.col, self.row = b[0], b[1], b[2], b[3], b[4], b[5]
        self.pixelSize = [b[6], b[7]]
        self.img00, self.img10, self.img01, self.img11 = b.vec3(8), b.vec3(11), b.vec3(14), b.vec3(17)
        self.rowWeight0, self.rowWeight1, self.normalization = b[20], b[21], b[22]


def env_sample_direction(e, r0, r1):     # src/envlight.cpp:291-322
    u0 = (r0 - e.cdfCol0) / (e.cdfCol1 - e.cdfCol0)
    u1 = (r1 - e.cdfRow0) / (e.cdfRow1 - e.cdfRow0)
    tx = tent(u0)
    ty = tent(u1)
    plx = e.col + tx
    ply = e.row + ty
    phi = (plx + 0.5) * e.pixelSize[0]
    theta = (ply + 0.5) * e.pixelSize[1]
    sinPhi = sin(phi)
    cosPhi = cos(phi)
    sinTheta = sin(theta)
    cosTheta = cos(theta)
    dirToLight = xform_vector(e.toWorld, [sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta])
    dx1, dx2, dy1, dy2 = tx, 1.0 - tx, ty, 1.0 - ty
    value1 = vadd(vmuls(vmuls(e.img00, dx2), dy2), vmuls(vmuls(e.img10, dx1), dy2))
    value2 = vadd(vmuls(vmuls(e.img01, dx2), dy1), vmuls(vmuls(e.img11, dx1), dy1))
    value = vadd(value1, value2)
    pdf = (luminance(value1) * e.rowWeight0 + luminance(value2) * e.rowWeight1) * e.normalization / \
        fmax(fabs(sinTheta), 1e-7)
    return dirToLight, value, pdf


def sample_direct(buffer, scn, pos, r0, r1):             # src/light.cpp SampleDirect
    t = buffer[0]
    ret = begin_if(Eq(t, LIGHT_POINT), 9)
    lightPos, emission = buffer.vec3(1), buffer.vec3(4)
    d = vsub(lightPos, pos)
    distSq = length_squared(d)
    directPdf = distSq
    dist = sqrt(distSq)
    d = [d[i] / dist for i in range(3)]
    lightContrib = vmuls(emission, inverse(distSq))
    set_cond_output(d + lightContrib + [C(1.0), directPdf, C(INVFOURPI)])
    begin_else_if(Eq(t, LIGHT_AREA))
    posOnLight, normalOnLight, shapePdf = sample_shape(buffer + 1, r0, r1)
    emission = (buffer + 1 + SER_SHAPE).vec3(0)
    d = vsub(posOnLight, pos)
    distSq = length_squared(d)
    dist = sqrt(distSq)
    d = [d[i] / dist for i in range(3)]
    cosAtLight = -dot(d, normalOnLight)
    directPdf = shapePdf * distSq / cosAtLight
    lightContrib = [emission[i] / directPdf for i in range(3)]
    emissionPdf = shapePdf * cosAtLight * INVPI
    set_cond_output(d + lightContrib + [cosAtLight, directPdf, emissionPdf])
    begin_else_if(Eq(t, LIGHT_ENV))
    e = EnvRec(buffer)
    d, value, directPdf = env_sample_direction(e, r0, r1)
    lightContrib = vmuls(value, inverse(directPdf))
    positionPdf = INVPI / square(scn.bsphereRadius)
    emissionPdf = directPdf * positionPdf
    set_cond_output(d + lightContrib + [C(1.0), directPdf, emissionPdf])
    begin_else()
    set_cond_output([C(0.0)] * 9)
    end_if()
    return ret[0:3], ret[3:6], ret[6], ret[7], ret[8]


def emission(buffer, scn, dirToLight, normalOnLight):    # src/light.cpp Emission
    t = buffer[0]
    ret = begin_if(Eq(t, LIGHT_AREA), 5)
    shapePdf = sample_shape_pdf(buffer + 1)
    em = (buffer + 1 + SER_SHAPE).vec3(0)
    cosAtLight = -dot(normalOnLight, dirToLight)
    directPdf = shapePdf
    emissionPdf = cosAtLight * directPdf * INVPI
    set_cond_output(em + [directPdf, emissionPdf])
    begin_else_if(Eq(t, LIGHT_ENV))
    e = EnvRec(buffer)
    d = xform_vector(e.toLight, dirToLight)
    uvx = atan2(d[0], -d[2]) / e.pixelSize[0] - 0.5
    uvy = acos(d[1]) / e.pixelSize[1] - 0.5
    dx1 = uvx - e.col
    dx2 = 1.0 - dx1
    dy1 = uvy - e.row
    dy2 = 1.0 - dy1
    value1 = vadd(vmuls(vmuls(e.img00, dx2), dy2), vmuls(vmuls(e.img10, dx1), dy2))
    value2 = vadd(vmuls(vmuls(e.img01, dx2), dy1), vmuls(vmuls(e.img11, dx1), dy1))
    em = vadd(value1, value2)
    sinTheta = sqrt(fmax(1.0 - square(d[1]), 1e-6))
    directPdf = (luminance(value1) * e.rowWeight0 + luminance(value2) * e.rowWeight1) * e.normalization / \
        fmax(fabs(sinTheta), 1e-7)
    positionPdf = INVPI / square(scn.bsphereRadius)
    emissionPdf = directPdf * positionPdf
    set_cond_output(em + [directPdf, emissionPdf])
    begin_else()
    set_cond_output([C(0.0)] * 5)
    end_if()
    return ret[0:3], ret[3], ret[4]


def emit(buffer, scn, p0, p1, d0, d1):                   # src/light.cpp Emit
    t = buffer[0]
    ret = begin_if(Eq(t, LIGHT_POINT), 12)
    lightPos, em = buffer.vec3(1), buffer.vec3(4)
    dirv, _ = sample_sphere(d0, d1)
    set_cond_output(lightPos + dirv + em + [C(1.0), C(INVFOURPI), C(1.0)])
    begin_else_if(Eq(t, LIGHT_AREA))
    org, normalOnLight, shapePdf = sample_shape(buffer + 1, p0, p1)
    em_ = (buffer + 1 + SER_SHAPE).vec3(0)
    d = sample_cos_hemisphere(d0, d1)
    b0, b1 = coordinate_system(normalOnLight)
    dirv = vadd(vadd(smulv(d[0], b0), smulv(d[1], b1)), smulv(d[2], normalOnLight))
    em = vmuls(em_, math.pi / shapePdf)
    cosAtLight = d[2]
    emissionPdf = d[2] * INVPI * shapePdf
    directPdf = shapePdf
    set_cond_output(org + dirv + em + [cosAtLight, emissionPdf, directPdf])
    begin_else_if(Eq(t, LIGHT_ENV))
    e = EnvRec(buffer)
    dirv, em, directPdf = env_sample_direction(e, d0, d1)
    dirv = vneg(dirv)
    ox, oy = sample_concentric_disc(p0, p1)
    b0, b1 = coordinate_system(dirv)
    perpOffset = vadd(smulv(ox, b0), smulv(oy, b1))
    org = vadd(scn.bsphereCenter, vmuls(vsub(perpOffset, dirv), scn.bsphereRadius))
    positionPdf = INVPI / square(scn.bsphereRadius)
    emissionPdf = directPdf * positionPdf
    set_cond_output(org + dirv + em + [C(1.0), emissionPdf, directPdf])
    begin_else()
    set_cond_output([C(0.0)] * 12)
    end_if()
    return ret[0:3], ret[3:6], ret[6:9], ret[9], ret[10], ret[11]


# ---- path stages (src/path.cpp:2799-3380) -----------------------------------------------------------
def emit_from_light(buffer, scn, lightPickProb, p0, p1, d0, d1, ray, ps):
    lightType = buffer[0]
    ray.org, ray.dir, ps.throughput, cosLight, emissionPdf, directPdf = emit(buffer, scn, p0, p1, d0, d1)
    emissionPdf = emissionPdf * lightPickProb
    directPdf = directPdf * lightPickProb
    ps.throughput = vmuls(ps.throughput, inverse(lightPickProb))
    ps.accMISWPrev = MIS(directPdf / emissionPdf)
    ret = begin_if(Eq(lightType, LIGHT_POINT), 1)       # sic: SURVEY.md App. B#4
    set_cond_output([MIS(cosLight / emissionPdf)])
    begin_else()
    set_cond_output([C(0.0)])
    end_if()
    ps.accMISWThis = ret[0]
    return buffer + SER_LIGHT


def convert_mis_light_emit(lightType, ray, ps):
    ret = begin_if(Eq(lightType, LIGHT_ENV), 1)
    set_cond_output([C(1.0)])
    begin_else()
    set_cond_output([MIS(distance_squared(ray.org, ps.position))])
    end_if()
    invCosTheta = inverse(MIS(fabs(dot(ray.dir, ps.shadingNormal))))
    ps.accMISWPrev = ps.accMISWPrev * (invCosTheta * ret[0])
    ps.accMISWThis = ps.accMISWThis * invCosTheta


def convert_mis_light_hit(lightType, ray, ps):
    ret = begin_if(Eq(lightType, LIGHT_ENV), 2)
    set_cond_output([C(1.0), C(1.0)])
    begin_else()
    distSq = MIS(distance_squared(ray.org, ps.position))
    invCosTheta = inverse(MIS(fabs(dot(ray.dir, ps.shadingNormal))))
    set_cond_output([invCosTheta, distSq])
    end_if()
    ps.accMISWPrev = ps.accMISWPrev * (ret[0] * ret[1])
    ps.accMISWThis = ps.accMISWThis * ret[0]


def convert_mis(ray, ps):
    ps.accMISWPrev = ps.accMISWPrev * MIS(distance_squared(ray.org, ps.position))
    invCosTheta = inverse(MIS(fabs(dot(ray.dir, ps.shadingNormal))))
    ps.accMISWPrev = ps.accMISWPrev * invCosTheta
    ps.accMISWThis = ps.accMISWThis * invCosTheta


def connect_to_camera(scn, buffer, ps):
    camOrg, camDir = sample_primary(scn, C(0.5), C(0.5))
    dirToCamera = vsub(camOrg, ps.position)
    distSq = length_squared(dirToCamera)
    dist = sqrt(distSq)
    dirToCamera = vmuls(dirToCamera, inverse(dist))
    bsdfContrib, cosToCamera, bsdfPdf, bsdfRevPdf = evaluate_bsdf(True, buffer, ps.wi, ps.shadingNormal, dirToCamera)
    factor = shading_normal_correction_adjoint(ps.wi, ps, dirToCamera)
    bsdfContrib = vmuls(bsdfContrib, factor)
    invCosAtCamera = -inverse(dot(camDir, dirToCamera))
    imagePointToCameraDist = scn.camDist * invCosAtCamera
    imageToSolidAngleFactor = square(imagePointToCameraDist) * invCosAtCamera
    imageToSurfaceFactor = imageToSolidAngleFactor * fabs(cosToCamera) / distSq
    cameraPdf = imageToSurfaceFactor
    wLight = MIS(cameraPdf / scn.screenPixelCount) * (ps.accMISWPrev + ps.accMISWThis * MIS(bsdfRevPdf))
    misWeight = inverse(wLight + 1.0)
    surfaceToImageFactor = cosToCamera / imageToSurfaceFactor
    denom = scn.screenPixelCount * surfaceToImageFactor
    contrib = [(misWeight * bsdfContrib[i]) / denom for i in range(3)]
    ps.throughput = cwise(contrib, ps.throughput)


def bsdf_sampling(adjoint, buffer, r0, r1, bsdfDiscrete, useAbsoluteParam, ps):
    ret = begin_if(Eq(useAbsoluteParam, 0.0), 10)
    wo, c, cw, p, rp = sample_bsdf(adjoint, buffer, ps.wi, ps.shadingNormal, r0, r1, bsdfDiscrete)
    set_cond_output(wo + c + [cw, p, rp, C(1.0)])
    begin_else()
    wo, jac = sample_sphere(r0, r1)
    c, cw, p, rp = evaluate_bsdf(adjoint, buffer, ps.wi, ps.shadingNormal, wo)
    set_cond_output(wo + c + [cw, p, rp, jac])
    end_if()
    dirv, bsdfContrib, cosWo, bsdfPdf, bsdfRevPdf, jacobian = ret[0:3], ret[3:6], ret[6], ret[7], ret[8], ret[9]
    if adjoint:
        factor = shading_normal_correction_adjoint(ps.wi, ps, dirv)
        bsdfContrib = vmuls(bsdfContrib, factor)
    bsdfContrib = vmuls(bsdfContrib, jacobian)
    ps.accMISWThis = MIS(cosWo / bsdfPdf) * (ps.accMISWThis * MIS(bsdfRevPdf) + ps.accMISWPrev)
    ps.accMISWPrev = MIS(inverse(bsdfPdf))
    ps.throughput = cwise(ps.throughput, bsdfContrib)
    return buffer + SER_BSDF, dirv


def emit_from_camera(scn, sx, sy, ray, ps):
    camOrg, camDir = sample_primary(scn, C(0.5), C(0.5))
    ray.org, ray.dir = sample_primary(scn, sx, sy)
    cosAtCamera = dot(camDir, ray.dir)
    imagePointToCameraDist = scn.camDist / cosAtCamera
    imageToSolidAngleFactor = square(imagePointToCameraDist) / cosAtCamera
    cameraPdf = imageToSolidAngleFactor
    ps.throughput = [C(1.0), C(1.0), C(1.0)]
    ps.accMISWPrev = MIS(scn.screenPixelCount / cameraPdf)
    ps.accMISWThis = C(0.0)


def handle_hit_light(scn, buffer, dirv, ps):
    em, directPdf, emissionPdf = emission(buffer, scn, dirv, ps.shadingNormal)
    buffer = buffer + SER_LIGHT
    ps.throughput = cwise(ps.throughput, em)
    lightPickProb = buffer[0]
    directPdf = directPdf * lightPickProb
    emissionPdf = emissionPdf * lightPickProb
    wCamera = MIS(directPdf) * ps.accMISWPrev + MIS(emissionPdf) * ps.accMISWThis
    misWeight = inverse(1.0 + wCamera)
    ps.throughput = vmuls(ps.throughput, misWeight)


def direct_lighting(scn, buffer, ps, r0, r1):
    lightType = buffer[0]
    dirToLight, lightContrib, cosAtLight, directPdf, emissionPdf = sample_direct(buffer, scn, ps.position, r0, r1)
    buffer = buffer + SER_LIGHT
    bsdfContrib, cosToLight, bsdfPdf, bsdfRevPdf = evaluate_bsdf(False, buffer, ps.wi, ps.shadingNormal, dirToLight)
    buffer = buffer + SER_BSDF
    lightPickProb = buffer[0]
    ps.throughput = cwise(ps.throughput, bsdfContrib)
    ps.throughput = vmuls(cwise(ps.throughput, lightContrib), inverse(lightPickProb))
    ret = begin_if(Eq(lightType, LIGHT_POINT), 1)
    set_cond_output([C(0.0)])
    begin_else()
    set_cond_output([MIS(bsdfPdf / (lightPickProb * directPdf))])
    end_if()
    wLight = ret[0]
    wCamera = MIS(emissionPdf * cosToLight / (directPdf * cosAtLight)) * (ps.accMISWPrev + ps.accMISWThis * MIS(bsdfRevPdf))
    misWeight = inverse(wLight + 1.0 + wCamera)
    ps.throughput = vmuls(ps.throughput, misWeight)


def connect_vertex(buffer, lgtBSDFBuffer, lps, cps):
    dirToLight = vsub(lps.position, cps.position)
    distSq = length_squared(dirToLight)
    dist = sqrt(distSq)
    dirToLight = vmuls(dirToLight, inverse(dist))
    camBsdfFactor, cosCamera, camBsdfPdf, camBsdfRevPdf = evaluate_bsdf(False, buffer, cps.wi, cps.shadingNormal, dirToLight)
    lgtBsdfFactor, cosLight, lgtBsdfPdf, lgtBsdfRevPdf = evaluate_bsdf(True, lgtBSDFBuffer, lps.wi, lps.shadingNormal,
                                                                       vneg(dirToLight))
    lgtFactor = shading_normal_correction_adjoint(lps.wi, lps, vneg(dirToLight))
    lgtBsdfFactor = vmuls(lgtBsdfFactor, lgtFactor)
    geometryTerm = inverse(distSq)
    camBsdfDirPdfA = camBsdfPdf * cosLight * geometryTerm
    lgtBsdfDirPdfA = lgtBsdfPdf * cosCamera * geometryTerm
    wLight = MIS(camBsdfDirPdfA) * (lps.accMISWPrev + lps.accMISWThis * MIS(lgtBsdfRevPdf))
    wCamera = MIS(lgtBsdfDirPdfA) * (cps.accMISWPrev + cps.accMISWThis * MIS(camBsdfRevPdf))
    misWeight = inverse(wLight + 1.0 + wCamera)
    cps.throughput = cwise(lps.throughput, cps.throughput)
    cps.throughput = cwise(cps.throughput, camBsdfFactor)
    cps.throughput = vmuls(cwise(cps.throughput, lgtBsdfFactor), geometryTerm * misWeight)


# ---- the whole function of a class (tests) ------------------------------------------------------------------
def record_path_function(maxCamDepth, maxLightDepth, params, pss):
    """pss: list of D differentiable input nodes (primary[1..D]).  Returns the logLum node."""
    scn = Scene(Buf(params, "scene"))
    buffer = Buf(params, "vert") + 3
    pi = 0
    lgtBSDFBuffer = None
    lps = PathState()
    contrib = [C(0.0)] * 3
    if maxLightDepth > 1:
        lightPickProb = buffer[0]
        buffer = buffer + 1
        ray = Ray()
        p0, p1, d0, d1 = pss[pi:pi + 4]
        pi += 4
        lightType = buffer[0]
        buffer = emit_from_light(buffer, scn, lightPickProb, p0, p1, d0, d1, ray, lps)
        for lgtDepth in range(maxLightDepth - 1):
            buffer = intersect(buffer, ray, lps)
            bsdfDiscrete, useAbsoluteParam = buffer[0], buffer[1]
            buffer = buffer + 2
            lps.wi = vneg(ray.dir)
            if lgtDepth == 0:
                convert_mis_light_emit(lightType, ray, lps)
            else:
                convert_mis(ray, lps)
            if lgtDepth == maxLightDepth - 2:
                if maxCamDepth == 1:
                    connect_to_camera(scn, buffer, lps)
                    contrib = lps.throughput
                lgtBSDFBuffer = buffer
                buffer = buffer + SER_BSDF
                break
            r0, r1 = pss[pi:pi + 2]
            pi += 2
            buffer, ray.dir = bsdf_sampling(True, buffer, r0, r1, bsdfDiscrete, useAbsoluteParam, lps)
            rrWeight = buffer[0]
            buffer = buffer + 1
            lps.throughput = vmuls(lps.throughput, rrWeight)
            ray.org = lps.position
    if maxCamDepth > 1:
        sx, sy = pss[pi:pi + 2]
        pi += 2
        ray = Ray()
        cps = PathState()
        emit_from_camera(scn, sx, sy, ray, cps)
        for camDepth in range(maxCamDepth - 1):
            buffer = intersect(buffer, ray, cps)
            cps.wi = vneg(ray.dir)
            if camDepth == maxCamDepth - 2 and maxLightDepth == 0:
                lightType = buffer[0]
                convert_mis_light_hit(lightType, ray, cps)
                handle_hit_light(scn, buffer, ray.dir, cps)
                contrib = cps.throughput
                break
            convert_mis(ray, cps)
            if camDepth == maxCamDepth - 2:
                if maxLightDepth == 1:
                    r0, r1 = pss[pi:pi + 2]
                    pi += 2
                    direct_lighting(scn, buffer, cps, r0, r1)
                else:
                    connect_vertex(buffer, lgtBSDFBuffer, lps, cps)
                contrib = cps.throughput
                break
            r0, r1 = pss[pi:pi + 2]
            pi += 2
            bsdfDiscrete, useAbsoluteParam = buffer[0], buffer[1]
            buffer = buffer + 2
            buffer, ray.dir = bsdf_sampling(False, buffer, r0, r1, bsdfDiscrete, useAbsoluteParam, cps)
            rrWeight = buffer[0]
            buffer = buffer + 1
            cps.throughput = vmuls(cps.throughput, rrWeight)
            ray.org = cps.position
    return log(luminance(contrib))
