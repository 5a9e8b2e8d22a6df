// host_scene.cpp -- host-side scene loader, flattener and SAH BVH2 builder.
//
// Re-states (behaviour only; written from scratch) the reference's
//   ParseScene & friends       src/parsescene.cpp:46-639
//   LoadSerialized             src/loadserialized.cpp:104-325
//   ParseObj                   src/parseobj.cpp:37-275
//   Transform helpers          src/transform.cpp:4-104, src/animatedtransform.cpp:10-46,
//                              src/quaternion.cpp:4-36, src/quaternion.h:13-38
//   Camera::Camera             src/camera.cpp:11-28
//   Scene::Scene               src/scene.cpp:8-46  (light cdf, bounding sphere x1000)
//   CreateEnvmapSampleInfo     src/envlight.cpp:24-71
//   TriangleMesh::SetAreaLight src/trianglemesh.cpp:293-307
//   Phong::GetKsWeight         src/phong.cpp:159-169, BitmapTexture::ComputeAvg bitmaptexture.h:99-133
//   PiecewiseConstant1D        src/distribution.h:9-29
// Embree's BVH build (src/trianglemesh.cpp:107-143) is replaced by the binned-SAH BVH2 below.
#include "host_scene.h"
#include "mini_xml.h"
#include "image_decode.h"
#include "../core/bsdf.h"
#include "../core/bvh.h"

#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstddef>
#include <fstream>
#include <sstream>

using namespace lmc;

namespace lmc_host {

// ------------------------------------------------------------------------------------------
// small matrix toolbox (row-major float, double internally only for the general inverse)
// ------------------------------------------------------------------------------------------
static M44 ident() {
    M44 m;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.m[i][j] = (i == j) ? 1.0f : 0.0f;
    return m;
}
static M44 mul(const M44 &a, const M44 &b) {
    M44 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
static M44 transpose(const M44 &a) {
    M44 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i];
    return r;
}
// General inverse by Gauss-Jordan in double (the reference calls Eigen's Matrix4f::inverse();
// pinned by our oracle only, tolerance 1e-6).
static M44 inverse44(const M44 &a) {
    double w[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { w[i][j] = a.m[i][j]; w[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++) if (fabs(w[r][c]) > fabs(w[piv][c])) piv = r;
        if (w[piv][c] == 0.0) throw std::runtime_error("singular matrix in scene description");
        if (piv != c) for (int j = 0; j < 8; j++) std::swap(w[piv][j], w[c][j]);
        const double inv = 1.0 / w[c][c];
        for (int j = 0; j < 8; j++) w[c][j] *= inv;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f != 0.0) for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j];
        }
    }
    M44 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = (float)w[i][4 + j];
    return r;
}
static M44 scale_m(float x, float y, float z) { M44 m = ident(); m.m[0][0] = x; m.m[1][1] = y; m.m[2][2] = z; return m; }
static M44 translate_m(float x, float y, float z) { M44 m = ident(); m.m[0][3] = x; m.m[1][3] = y; m.m[2][3] = z; return m; }
static float radians(float d) { return d * (LMC_PI / 180.0f); }
static V3 hnormalize(V3 v) { const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); return mk3(v.x / l * 1.0f, v.y / l, v.z / l); }
static M44 rotate_m(float angle, V3 axis) {
    const V3 a = normalize(axis);
    const float s = sinf(radians(angle)), c = cosf(radians(angle));
    M44 m = ident();
    m.m[0][0] = a.x * a.x + (1.0f - a.x * a.x) * c;
    m.m[0][1] = a.x * a.y * (1.0f - c) - a.z * s;
    m.m[0][2] = a.x * a.z * (1.0f - c) + a.y * s;
    m.m[1][0] = a.x * a.y * (1.0f - c) + a.z * s;
    m.m[1][1] = a.y * a.y + (1.0f - a.y * a.y) * c;
    m.m[1][2] = a.y * a.z * (1.0f - c) - a.x * s;
    m.m[2][0] = a.x * a.z * (1.0f - c) - a.y * s;
    m.m[2][1] = a.y * a.z * (1.0f - c) + a.x * s;
    m.m[2][2] = a.z * a.z + (1.0f - a.z * a.z) * c;
    return m;
}
static M44 lookat_m(V3 pos, V3 look, V3 up) {
    const V3 dir = normalize(look - pos);
    const V3 cr = cross(normalize(up), dir);
    if (length(cr) == 0.0f) throw std::runtime_error("lookat: up vector parallel to viewing direction");
    const V3 left = normalize(cr);
    const V3 newUp = cross(dir, left);
    M44 m = ident();
    m.m[0][0] = left.x; m.m[1][0] = left.y; m.m[2][0] = left.z;
    m.m[0][1] = newUp.x; m.m[1][1] = newUp.y; m.m[2][1] = newUp.z;
    m.m[0][2] = dir.x; m.m[1][2] = dir.y; m.m[2][2] = dir.z;
    m.m[0][3] = pos.x; m.m[1][3] = pos.y; m.m[2][3] = pos.z;
    return m;
}
static M44 perspective_m(float fov, float clipNear, float clipFar) {
    const float recip = 1.0f / (clipFar - clipNear);
    const float cot = 1.0f / tanf(radians(fov / 2.0f));
    M44 m;
    memset(&m, 0, sizeof(m));
    m.m[0][0] = cot; m.m[1][1] = cot; m.m[2][2] = clipFar * recip; m.m[2][3] = -clipNear * clipFar * recip;
    m.m[3][2] = 1.0f;
    return m;
}

// rigid transform -> (translate, quaternion): the reference's Decompose runs a JacobiSVD polar
// decomposition first; for a pure rotation the polar factor is the matrix itself.
struct Rigid { float t[3]; float q[4]; };
static Rigid decompose(const M44 &m) {
    // verify there is no scaling (the reference throws "Scaling in animation")
    for (int c = 0; c < 3; c++) {
        const float l2 = m.m[0][c] * m.m[0][c] + m.m[1][c] * m.m[1][c] + m.m[2][c] * m.m[2][c];
        if (fabsf(l2 - 1.0f) > 1e-4f) throw std::runtime_error("camera/envmap transform must be rigid (no scaling)");
    }
    Rigid r;
    float q[4];
    const float trace = m.m[0][0] + m.m[1][1] + m.m[2][2];
    if (trace > 1e-7f) {
        float s = sqrtf(trace + 1.0f);
        q[3] = s / 2.0f;
        s = 0.5f / s;
        q[0] = (m.m[2][1] - m.m[1][2]) * s;
        q[1] = (m.m[0][2] - m.m[2][0]) * s;
        q[2] = (m.m[1][0] - m.m[0][1]) * s;
    } else {
        const int nxt[3] = {1, 2, 0};
        float _q[3];
        int i = 0;
        if (m.m[1][1] > m.m[0][0]) i = 1;
        if (m.m[2][2] > m.m[i][i]) i = 2;
        const int j = nxt[i], k = nxt[j];
        float s = sqrtf((m.m[i][i] - (m.m[j][j] + m.m[k][k])) + 1.0f);
        _q[i] = s * 0.5f;
        if (s != 0.0f) s = 0.5f / s;
        q[3] = (m.m[k][j] - m.m[j][k]) * s;
        _q[j] = (m.m[j][i] + m.m[i][j]) * s;
        _q[k] = (m.m[k][i] + m.m[i][k]) * s;
        q[0] = _q[0]; q[1] = _q[1]; q[2] = _q[2];
    }
    for (int i = 0; i < 4; i++) r.q[i] = q[i];
    r.t[0] = m.m[0][3]; r.t[1] = m.m[1][3]; r.t[2] = m.m[2][3];
    return r;
}
static M44 quat_to_m(const float *q) {
    const float xx = q[0] * q[0], yy = q[1] * q[1], zz = q[2] * q[2];
    const float xy = q[0] * q[1], xz = q[0] * q[2], yz = q[1] * q[2];
    const float wx = q[0] * q[3], wy = q[1] * q[3], wz = q[2] * q[3];
    M44 m = ident();
    m.m[0][0] = 1.0f - 2.0f * (yy + zz); m.m[0][1] = 2.0f * (xy + wz); m.m[0][2] = 2.0f * (xz - wy);
    m.m[1][0] = 2.0f * (xy - wz); m.m[1][1] = 1.0f - 2.0f * (xx + zz); m.m[1][2] = 2.0f * (yz + wx);
    m.m[2][0] = 2.0f * (xz + wy); m.m[2][1] = 2.0f * (yz - wx); m.m[2][2] = 1.0f - 2.0f * (xx + yy);
    return transpose(m);
}
static Rigid invert_rigid(const Rigid &r) {
    Rigid o;
    o.q[0] = -r.q[0]; o.q[1] = -r.q[1]; o.q[2] = -r.q[2]; o.q[3] = r.q[3];
    const M44 rot = quat_to_m(o.q);
    const V3 t = xform_vector(rot, mk3(r.t[0], r.t[1], r.t[2]));
    o.t[0] = -t.x; o.t[1] = -t.y; o.t[2] = -t.z;
    return o;
}
static M44 rigid_to_m(const Rigid &r) { return mul(translate_m(r.t[0], r.t[1], r.t[2]), quat_to_m(r.q)); }
static void rigid_serialize(const Rigid &r, float *out15) {
    out15[0] = 0.0f;  // isMoving
    for (int i = 0; i < 3; i++) { out15[1 + i] = r.t[i]; out15[4 + i] = r.t[i]; }
    for (int i = 0; i < 4; i++) { out15[7 + i] = r.q[i]; out15[11 + i] = r.q[i]; }
}

// ------------------------------------------------------------------------------------------
// parsing helpers
// ------------------------------------------------------------------------------------------
static std::vector<std::string> split_list(const std::string &value) {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : value) {
        if (ch == ',' || ch == ' ' || ch == '\t' || ch == '\n') {
            if (!cur.empty()) { out.push_back(cur); cur.clear(); }
        } else cur.push_back(ch);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}
static V3 parse_vector3(const std::string &value) {
    const auto l = split_list(value);
    if (l.size() == 1) { const float f = std::stof(l[0]); return mk3(f, f, f); }
    if (l.size() == 3) return mk3(std::stof(l[0]), std::stof(l[1]), std::stof(l[2]));
    throw std::runtime_error("ParseVector3 failed: '" + value + "'");
}
static float attr_f(const XmlNode &n, const char *k, float def) { return n.has(k) ? std::stof(n.attr(k)) : def; }
static std::string lower(std::string s) { for (auto &c : s) c = (char)tolower((unsigned char)c); return s; }

static M44 parse_transform(const XmlNode &node) {
    M44 t = ident();
    for (auto &cp : node.children) {
        const XmlNode &c = *cp;
        const std::string name = lower(c.name);
        if (name == "scale") {
            if (c.has("value")) { const float s = std::stof(c.attr("value")); t = mul(scale_m(s, s, s), t); }
            else t = mul(scale_m(attr_f(c, "x", 1), attr_f(c, "y", 1), attr_f(c, "z", 1)), t);
        } else if (name == "translate") {
            t = mul(translate_m(attr_f(c, "x", 0), attr_f(c, "y", 0), attr_f(c, "z", 0)), t);
        } else if (name == "rotate") {
            t = mul(rotate_m(attr_f(c, "angle", 0), mk3(attr_f(c, "x", 0), attr_f(c, "y", 0), attr_f(c, "z", 0))), t);
        } else if (name == "lookat") {
            t = mul(lookat_m(parse_vector3(c.attr("origin")), parse_vector3(c.attr("target")), parse_vector3(c.attr("up"))), t);
        } else if (name == "matrix") {
            const auto l = split_list(c.attr("value"));
            if (l.size() != 16) throw std::runtime_error("ParseMatrix4x4 failed");
            M44 m;
            for (int i = 0; i < 16; i++) m.m[i / 4][i % 4] = std::stof(l[i]);
            t = mul(m, t);
        }
    }
    return t;
}

// ------------------------------------------------------------------------------------------
// images: PNG / JPEG / OpenEXR are decoded here (image_decode.h, jpeg_decode.h; what the reference does through
// OpenImageIO, src/bitmaptexture.h:73-146, src/image.cpp:5-45).  Any other format can be handed over pre-decoded as
// "<file>.rawf" (tools/stage_scenes.py; the containers of the bundled images are kept under tests/golden/decoded/ as the
// OpenCV-decoded ground truth the native decoders are tested against).
// ------------------------------------------------------------------------------------------
typedef DecodedImage RawImage;
static RawImage load_rawf(const std::string &path) {
    {
        std::ifstream probe(path, std::ios::binary);
        RawImage native;
        if (probe && decode_image_native(path, native)) return native;
    }
    std::ifstream f(path + ".rawf", std::ios::binary);
    if (!f) throw std::runtime_error("cannot open image '" + path + "' (PNG / JPEG / OpenEXR are decoded natively; other formats: '" + path + ".rawf', see tools/stage_scenes.py)");
    char magic[4]; int hdr[3];
    f.read(magic, 4); f.read((char *)hdr, 12);
    if (memcmp(magic, "RAWF", 4) != 0) throw std::runtime_error("bad rawf magic: " + path);
    RawImage im; im.w = hdr[0]; im.h = hdr[1]; im.is8 = hdr[2];
    const size_t n = (size_t)im.w * im.h * 3;
    im.rgb.resize(n);
    if (im.is8) {
        std::vector<unsigned char> b(n);
        f.read((char *)b.data(), n);
        for (size_t i = 0; i < n; i++) im.rgb[i] = (float)b[i] / 255.0f;
    } else {
        f.read((char *)im.rgb.data(), n * 4);
    }
    if (!f) throw std::runtime_error("short read: " + path);
    return im;
}

// ------------------------------------------------------------------------------------------
// meshes
// ------------------------------------------------------------------------------------------
struct Mesh {
    std::vector<V3> pos, nor;
    std::vector<V2> st;
    std::vector<uint32_t> idx;   // 3 per triangle
};

static float unit_angle(V3 u, V3 v) {
    if (dot(u, v) < 0.0f) return (LMC_PI - 2.0f) * asinf(0.5f * length(v + u));
    return 2.0f * asinf(0.5f * length(v - u));
}
// flipMode: 0 none, 1 = loadserialized.cpp's in-loop flip (src/loadserialized.cpp:136-137)
static void compute_normals(const Mesh &m, std::vector<V3> &normals, bool flipInLoop) {
    normals.assign(m.pos.size(), mk3s(0.0f));
    for (size_t t = 0; t < m.idx.size() / 3; t++) {
        V3 n = mk3s(0.0f);
        for (int i = 0; i < 3; ++i) {
            const uint32_t i0 = m.idx[3 * t + i], i1 = m.idx[3 * t + (i + 1) % 3], i2 = m.idx[3 * t + (i + 2) % 3];
            const V3 sideA = m.pos[i1] - m.pos[i0], sideB = m.pos[i2] - m.pos[i0];
            if (i == 0) {
                n = cross(sideA, sideB);
                const float len = length(n);
                if (len == 0.0f) break;
                n = n / len;
            }
            const float angle = unit_angle(normalize(sideA), normalize(sideB));
            normals[i0] = normals[i0] + n * angle;
            if (flipInLoop) normals[i0] = -normals[i0];
        }
    }
    for (auto &n : normals) {
        const float len = length(n);
        if (len != 0.0f) n = n / len; else n = mk3s(0.0f);
    }
}

static std::vector<unsigned char> inflate_all(const unsigned char *src, size_t n) {
    z_stream zs; memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15) != Z_OK) throw std::runtime_error("inflateInit2 failed");
    std::vector<unsigned char> out; out.resize(n * 4 + 65536);
    zs.next_in = (Bytef *)src; zs.avail_in = (uInt)n;
    size_t produced = 0;
    for (;;) {
        if (produced == out.size()) out.resize(out.size() * 2);
        zs.next_out = out.data() + produced; zs.avail_out = (uInt)(out.size() - produced);
        const int rc = inflate(&zs, Z_NO_FLUSH);
        produced = out.size() - zs.avail_out;
        if (rc == Z_STREAM_END) break;
        if (rc != Z_OK) { inflateEnd(&zs); throw std::runtime_error("inflate failed in .serialized mesh"); }
    }
    inflateEnd(&zs);
    out.resize(produced);
    return out;
}

static Mesh load_serialized(const std::string &filename, int shapeIndex, const M44 &toWorld, bool flipNormals,
                            bool faceNormals) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + filename);
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (file.size() < 8) throw std::runtime_error("truncated .serialized file");
    uint16_t version; memcpy(&version, &file[2], 2);
    size_t offset = 4;
    if (shapeIndex > 0) {
        uint32_t count; memcpy(&count, &file[file.size() - 4], 4);
        if ((uint32_t)shapeIndex >= count) throw std::runtime_error("shapeIndex out of range");
        if (version == 4) {
            uint64_t o; memcpy(&o, &file[file.size() - 4 - 8 * (count - shapeIndex)], 8); offset = (size_t)o;
        } else {
            uint32_t o; memcpy(&o, &file[file.size() - 4 * (count - shapeIndex + 1)], 4); offset = o;
        }
        offset += 4;  // skip the per-mesh header
    }
    const std::vector<unsigned char> z = inflate_all(&file[offset], file.size() - offset);
    size_t p = 0;
    auto rd = [&](void *dst, size_t n) {
        if (p + n > z.size()) throw std::runtime_error("short .serialized stream");
        memcpy(dst, &z[p], n); p += n;
    };
    uint32_t flags; rd(&flags, 4);
    if (version == 4) { char c; do { rd(&c, 1); } while (c != 0); }
    uint64_t vc, tc; rd(&vc, 8); rd(&tc, 8);
    const bool dbl = (flags & 0x2000) != 0;
    faceNormals = ((flags & 0x0010) != 0) || faceNormals;
    const M44 inv = inverse44(toWorld);
    Mesh m;
    auto rdf = [&]() -> float { if (dbl) { double d; rd(&d, 8); return (float)d; } float x; rd(&x, 4); return x; };
    m.pos.resize(vc);
    for (size_t i = 0; i < vc; i++) { const float x = rdf(), y = rdf(), zc = rdf(); m.pos[i] = xform_point(toWorld, mk3(x, y, zc)); }
    if (flags & 0x0001) {
        m.nor.resize(vc);
        for (size_t i = 0; i < vc; i++) {
            const float x = rdf(), y = rdf(), zc = rdf();
            V3 n = xform_vector_t(inv, mk3(x, y, zc));   // XformNormal(invXform, v)
            if (flipNormals) n = -n;
            m.nor[i] = n;
        }
    }
    if (flags & 0x0002) {
        m.st.resize(vc);
        for (size_t i = 0; i < vc; i++) { const float u = rdf(), v = rdf(); m.st[i] = mk2(u, v); }
    }
    if (flags & 0x0008) p += vc * 24;  // colours: always 3 doubles, unused
    m.idx.resize(tc * 3);
    rd(m.idx.data(), tc * 12);
    if (m.nor.empty() || faceNormals) compute_normals(m, m.nor, flipNormals);
    return m;
}

static Mesh parse_obj(const std::string &filename, const M44 &toWorld, bool flipNormals, bool faceNormals) {
    std::ifstream ifs(filename);
    if (!ifs) throw std::runtime_error("Unable to open the obj file " + filename);
    std::vector<V3> posPool, norPool; std::vector<V2> stPool;
    struct Key { int v, vt, vn; bool operator<(const Key &o) const { if (v != o.v) return v < o.v; if (vt != o.vt) return vt < o.vt; return vn < o.vn; } };
    std::map<Key, uint32_t> vmap;
    Mesh m;
    const M44 inv = inverse44(toWorld);
    auto face_key = [](const std::string &s) {
        int r[3] = {0, 0, 0}; int k = 0; std::string cur;
        for (size_t i = 0; i <= s.size() && k < 3; i++) {
            if (i == s.size() || s[i] == '/') { r[k++] = cur.empty() ? 0 : std::stoi(cur); cur.clear(); }
            else cur.push_back(s[i]);
        }
        Key key; key.v = r[0] - 1; key.vt = r[1] - 1; key.vn = r[2] - 1; return key;
    };
    auto vid = [&](const Key &k) -> uint32_t {
        auto it = vmap.find(k);
        if (it != vmap.end()) return it->second;
        const uint32_t id = (uint32_t)m.pos.size();
        m.pos.push_back(xform_point(toWorld, posPool.at(k.v)));
        if (k.vt != -1) m.st.push_back(stPool.at(k.vt));
        if (k.vn != -1) m.nor.push_back(xform_vector_t(inv, norPool.at(k.vn)));
        vmap[k] = id;
        return id;
    };
    std::string line;
    while (std::getline(ifs, line)) {
        size_t b = 0; while (b < line.size() && isspace((unsigned char)line[b])) b++;
        if (b >= line.size() || line[b] == '#') continue;
        std::stringstream ss(line.substr(b));
        std::string tok; ss >> tok;
        if (tok == "v") { float x = 0, y = 0, z = 0, w = 1.f; ss >> x >> y >> z; if (!(ss >> w)) w = 1.f; const float iw = 1.f / w; posPool.push_back(mk3(x * iw, y * iw, z * iw)); }
        else if (tok == "vt") { float s = 0, t = 0; ss >> s >> t; stPool.push_back(mk2(s, 1.0f - t)); }
        else if (tok == "vn") { float x = 0, y = 0, z = 0; ss >> x >> y >> z; norPool.push_back(normalize(mk3(x, y, z))); }
        else if (tok == "f") {
            std::string i0, i1, i2, i3, i4; ss >> i0 >> i1 >> i2;
            const uint32_t a = vid(face_key(i0)), bb = vid(face_key(i1)), c = vid(face_key(i2));
            m.idx.push_back(a); m.idx.push_back(bb); m.idx.push_back(c);
            if (ss >> i3) {
                const uint32_t d = vid(face_key(i3));
                m.idx.push_back(a); m.idx.push_back(c); m.idx.push_back(d);
                if (ss >> i4) throw std::runtime_error("The object file contains n-gon (n>4) that we do not support.");
            }
        }
    }
    if (m.nor.empty() || faceNormals) compute_normals(m, m.nor, false);
    if (flipNormals) for (auto &n : m.nor) n = -n;
    if (!m.st.empty() && m.st.size() != m.pos.size()) throw std::runtime_error("obj: mixed vertices with/without texture coordinates: " + filename);
    if (m.nor.size() != m.pos.size()) throw std::runtime_error("obj: mixed vertices with/without normals: " + filename);
    return m;
}

// ------------------------------------------------------------------------------------------
// PiecewiseConstant1D cdf (src/distribution.h:9-29)
// ------------------------------------------------------------------------------------------
static std::vector<float> piecewise_cdf(const std::vector<float> &f) {
    const int n = (int)f.size();
    std::vector<float> cdf(n + 1);
    cdf[0] = 0.0f;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / n;
    const float funcInt = cdf[n];
    if (funcInt == 0.f) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    return cdf;
}

// ------------------------------------------------------------------------------------------
// SAH BVH2 builder (binned, 16 bins, leaves of <= 4 triangles)
// ------------------------------------------------------------------------------------------
struct BBoxF {
    float mn[3], mx[3];
    BBoxF() { for (int i = 0; i < 3; i++) { mn[i] = INFINITY; mx[i] = -INFINITY; } }
    void grow(const float *p) { for (int i = 0; i < 3; i++) { mn[i] = std::min(mn[i], p[i]); mx[i] = std::max(mx[i], p[i]); } }
    void merge(const BBoxF &b) { for (int i = 0; i < 3; i++) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); } }
    float area() const {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        if (dx < 0) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};
struct BuildPrim { BBoxF box; float c[3]; int src; };

struct Builder {
    std::vector<BuildPrim> prims;
    std::vector<BvhNode> nodes;
    std::vector<int> order;     // triangle order (src indices)
    static const int MAXLEAF = 4;

    int make_leaf(int lo, int hi) {
        const int first = (int)order.size();
        for (int i = lo; i < hi; i++) order.push_back(prims[i].src);
        return ~((first << 3) | (hi - lo - 1));
    }
    // returns child reference; bounds written to `box`
    // Depth budget.  The traversal keeps at most one deferred child per inner node on the way down, in a fixed
    // stack of LMC_BVH_STACK entries (core/bvh.h, cuda/trace_kernels.cuh), so no leaf may sit below LMC_BVH_STACK - 1
    // inner nodes.  A median split halves the primitive count, i.e. a subtree of n primitives built from median
    // splits alone is at most ceil(log2 n) deep; SAH splits (which may be arbitrarily unbalanced on skewed
    // geometry) are therefore only taken while depth + ceil(log2 n) leaves one level of slack.  maxDepth records
    // what was built and is checked again after the build and when a .pack is loaded (bvh_depth()).
    int maxDepth = 0;
    static int ceil_log2(int n) { int b = 0; while ((1 << b) < n) b++; return b; }
    int build(int lo, int hi, BBoxF &box, int depth) {
        box = BBoxF();
        BBoxF cb;
        for (int i = lo; i < hi; i++) { box.merge(prims[i].box); cb.grow(prims[i].c); }
        const int n = hi - lo;
        if (depth > maxDepth) maxDepth = depth;
        const bool sahAllowed = depth + ceil_log2(n) <= LMC_BVH_STACK - 2;
        if (n == 1 || (!sahAllowed && n <= 8 && depth + 1 <= LMC_BVH_STACK - 1)) return make_leaf(lo, hi);
        int mid = -1;
        if (sahAllowed) {
            // binned SAH on the widest centroid axis, falling back over all axes
            float bestCost = INFINITY; int bestAxis = -1, bestSplit = -1;
            const int NB = 16;
            for (int axis = 0; axis < 3; axis++) {
                const float ext = cb.mx[axis] - cb.mn[axis];
                if (!(ext > 0.0f)) continue;
                BBoxF bb[NB]; int cnt[NB] = {0};
                const float k = NB * (1.0f - 1e-6f) / ext;
                for (int i = lo; i < hi; i++) {
                    int b = (int)((prims[i].c[axis] - cb.mn[axis]) * k);
                    b = std::max(0, std::min(NB - 1, b));
                    bb[b].merge(prims[i].box); cnt[b]++;
                }
                float rightArea[NB]; int rightCnt[NB];
                BBoxF acc; int c = 0;
                for (int b = NB - 1; b > 0; b--) { acc.merge(bb[b]); c += cnt[b]; rightArea[b] = acc.area(); rightCnt[b] = c; }
                acc = BBoxF(); c = 0;
                for (int b = 0; b < NB - 1; b++) {
                    acc.merge(bb[b]); c += cnt[b];
                    if (c == 0 || rightCnt[b + 1] == 0) continue;
                    const float cost = acc.area() * c + rightArea[b + 1] * rightCnt[b + 1];
                    if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = b; }
                }
            }
            if (bestAxis >= 0) {
                const float leafCost = box.area() * n;
                if (n <= MAXLEAF && bestCost + box.area() >= leafCost) return make_leaf(lo, hi);   // Ctrav = Cisect
                const float ext = cb.mx[bestAxis] - cb.mn[bestAxis];
                const float k = NB * (1.0f - 1e-6f) / ext;
                auto it = std::partition(prims.begin() + lo, prims.begin() + hi, [&](const BuildPrim &p) {
                    int b = (int)((p.c[bestAxis] - cb.mn[bestAxis]) * k);
                    b = std::max(0, std::min(NB - 1, b));
                    return b <= bestSplit;
                });
                mid = (int)(it - prims.begin());
            }
        }
        if (mid <= lo || mid >= hi) {
            if (n <= 8) return make_leaf(lo, hi);
            // median split on the widest axis of the box
            int axis = 0; float e = -1;
            for (int a = 0; a < 3; a++) if (box.mx[a] - box.mn[a] > e) { e = box.mx[a] - box.mn[a]; axis = a; }
            mid = lo + n / 2;
            std::nth_element(prims.begin() + lo, prims.begin() + mid, prims.begin() + hi,
                             [axis](const BuildPrim &a, const BuildPrim &b) { return a.c[axis] < b.c[axis]; });
        }
        const int me = (int)nodes.size();
        nodes.push_back(BvhNode());
        BBoxF lb, rb;
        const int l = build(lo, mid, lb, depth + 1);
        const int r = build(mid, hi, rb, depth + 1);
        BvhNode &nd = nodes[me];
        for (int i = 0; i < 3; i++) { nd.lmin[i] = lb.mn[i]; nd.lmax[i] = lb.mx[i]; nd.rmin[i] = rb.mn[i]; nd.rmax[i] = rb.mx[i]; }
        nd.left = l; nd.right = r; nd.pad[0] = nd.pad[1] = 0;
        return me;
    }
};

// ------------------------------------------------------------------------------------------
// scene assembly
// ------------------------------------------------------------------------------------------
Options default_options() {
    Options o;
    memset(&o, 0, sizeof(o));
    o.minDepth = -1; o.maxDepth = 8; o.bidirectional = 1; o.h2mc = 0; o.mala = 0;
    o.numChains = 128; o.seedOffset = 0; o.useLightCoordinateSampling = 0; o.largeStepMultiplexed = 0;
    o.cacheEnabled = 0; o.maxDervDepth = 8; o.pssMinLength = 2; o.pssMaxLength = 12; o.adjointCompat = 1;
    o.outlierWeakRejectCnt = 10000; o.outlierStrongRejectCnt = 1000; o.outlierRatioThreshold = 30.0f;
    o.perturbStdDev = 0.01f; o.roughnessThreshold = 0.05f; o.largeStepProbability = 0.05f;
    o.largeStepProbScale = 1.0f; o.malaGN = 100.0f; o.malaStepsize = 0.005f; o.malaStdDev = 0.005f;
    o.discreteStdDev = 0.01f; o.uniformMixingProbability = 0.1f; o.lsRatio = 0.1f;
    // H2MCParam(sigma, L = Float(M_PI / 2.0)), src/h2mc.h:9-16 (float exp/sin/cos as in the reference)
    const float L = (float)(M_PI / 2.0);
    o.h2mcL = L;
    o.h2mcPosScale = 0.5f * (expf(L) - expf(-L)) * 0.5f * (expf(L) - expf(-L));
    o.h2mcPosOffset = 0.5f * (expf(L) + expf(-L) - 1.0f);
    o.h2mcNegScale = sinf(L) * sinf(L);
    o.h2mcNegOffset = -(cosf(L) - 1.0f);
    return o;
}

namespace {
struct OptField { const char *name; int isFloat; size_t off; };
#define LMC_OI(n, f) {n, 0, offsetof(Options, f)}
#define LMC_OF(n, f) {n, 1, offsetof(Options, f)}
const OptField kOptFields[] = {
    LMC_OI("mindepth", minDepth), LMC_OI("maxdepth", maxDepth), LMC_OI("bidirectional", bidirectional),
    LMC_OI("h2mc", h2mc), LMC_OI("mala", mala), LMC_OI("numchains", numChains), LMC_OI("seedoffset", seedOffset),
    LMC_OI("uselightcoordinatesampling", useLightCoordinateSampling), LMC_OI("largestepmultiplexed", largeStepMultiplexed),
    LMC_OI("maxdervdepth", maxDervDepth), LMC_OI("pssminlength", pssMinLength), LMC_OI("pssmaxlength", pssMaxLength),
    LMC_OI("adjointcompat", adjointCompat), LMC_OI("globalcache", cacheEnabled), LMC_OI("outlierweakrejectcnt", outlierWeakRejectCnt),
    LMC_OI("outlierstrongrejectcnt", outlierStrongRejectCnt), LMC_OF("outlierratiothreshold", outlierRatioThreshold),
    LMC_OF("perturbstddev", perturbStdDev), LMC_OF("roughnessthreshold", roughnessThreshold),
    LMC_OF("largestepprob", largeStepProbability), LMC_OF("largestepscale", largeStepProbScale),
    LMC_OF("mala-gn", malaGN), LMC_OF("mala-stepsize", malaStepsize), LMC_OF("malastddev", malaStdDev),
    LMC_OF("discretestddev", discreteStdDev), LMC_OF("uniformmixprob", uniformMixingProbability), LMC_OF("lsratio", lsRatio),
};
}  // namespace

bool set_option(Options &o, const std::string &name, double value) {
    for (const OptField &f : kOptFields) if (name == f.name) {
        if (f.isFloat) *(float *)((char *)&o + f.off) = (float)value; else *(int *)((char *)&o + f.off) = (int)value;
        return true;
    }
    return false;
}
bool get_option(const Options &o, const std::string &name, double &value) {
    for (const OptField &f : kOptFields) if (name == f.name) {
        value = f.isFloat ? (double)*(const float *)((const char *)&o + f.off) : (double)*(const int *)((const char *)&o + f.off);
        return true;
    }
    return false;
}

struct MatTex { std::string file; float s = 1, t = 1; };    // a bitmap-textured parameter (empty file = constant)
struct MatDesc {
    Material m;
    std::string kdTexFile; float sScale = 1, tScale = 1;   // Kd bitmap, if any
    MatTex ks, kt, exponent, alpha;
};

struct Loader {
    std::string baseDir;
    SceneStore &out;
    std::map<std::string, MatDesc> bsdfMap;
    struct TexDesc { std::string file; float s, t; };
    std::map<std::string, TexDesc> textureMap;
    std::map<std::string, int> texIndex;   // file|s|t -> index in out.textures
    std::map<std::string, RawImage> imageCache;

    explicit Loader(SceneStore &o) : out(o) {}

    const RawImage &image(const std::string &file) {
        auto it = imageCache.find(file);
        if (it != imageCache.end()) return it->second;
        return imageCache[file] = load_rawf(baseDir + file);
    }
    // BitmapTexture::ComputeAvg: mean of pow(pixel, gamma)
    V3 texture_avg(const std::string &file) {
        const RawImage &im = image(file);
        const float gamma = im.is8 ? 2.2f : 1.0f;
        double acc[3] = {0, 0, 0};
        const size_t n = (size_t)im.w * im.h;
        for (size_t i = 0; i < n; i++) for (int c = 0; c < 3; c++) acc[c] += powf(im.rgb[3 * i + c], gamma);
        return mk3((float)(acc[0] / n), (float)(acc[1] / n), (float)(acc[2] / n));
    }
    int texture_id(const TexDesc &t) {
        std::ostringstream key; key << t.file << "|" << t.s << "|" << t.t;
        auto it = texIndex.find(key.str());
        if (it != texIndex.end()) return it->second;
        const RawImage &im = image(t.file);
        Texture tx; tx.width = im.w; tx.height = im.h; tx.offset = (int)out.texData.size();
        tx.gamma = im.is8 ? 2.2f : 1.0f; tx.sScale = t.s; tx.tScale = t.t;
        out.texData.insert(out.texData.end(), im.rgb.begin(), im.rgb.end());
        out.textures.push_back(tx);
        return texIndex[key.str()] = (int)out.textures.size() - 1;
    }
    TexDesc parse_texture(const XmlNode &n) {
        if (n.attr("type") != "bitmap") throw std::runtime_error("Unknown texture type");
        TexDesc t; t.s = t.t = 1.0f;
        for (auto &c : n.children) {
            const std::string name = c->attr("name");
            if (name == "filename") t.file = c->attr("value");
            else if (name == "uvscale") t.s = t.t = std::stof(c->attr("value"));
        }
        return t;
    }
    // Parse3DMap: constant or texture (by <texture> child or <ref id>)
    void parse_rgb_map(const XmlNode &n, float *constant, bool &isTex, TexDesc &tex) {
        isTex = false;
        if (n.name == "texture") { tex = parse_texture(n); isTex = true; }
        else if (n.name == "ref") {
            auto it = textureMap.find(n.attr("id"));
            if (it == textureMap.end()) throw std::runtime_error("ref not found: " + n.attr("id"));
            tex = it->second; isTex = true;
        } else {
            const V3 v = parse_vector3(n.attr("value"));
            constant[0] = v.x; constant[1] = v.y; constant[2] = v.z;
        }
    }
    // Parse1DMap: a float constant or a bitmap whose first channel is used
    void parse_1d_map(const XmlNode &n, float &constant, MatTex &mt) {
        float dummy[3]; bool isTex; TexDesc tex;
        if (n.name == "texture" || n.name == "ref") {
            parse_rgb_map(n, dummy, isTex, tex);
            mt.file = tex.file; mt.s = tex.s; mt.t = tex.t;
        } else {
            constant = std::stof(n.attr("value"));
        }
    }
    MatDesc parse_bsdf(const XmlNode &n, bool twoSided) {
        const std::string type = n.attr("type");
        MatDesc d; memset(&d.m, 0, sizeof(d.m));
        d.m.twoSided = twoSided ? 1 : 0; d.m.kdTex = -1; d.m.areaLight = -1;
        d.m.ksTex = d.m.ktTex = d.m.expTex = d.m.alphaTex = -1;
        bool isTex; TexDesc tex;
        if (type == "diffuse") {
            d.m.type = BSDF_LAMBERTIAN;
            d.m.Kd[0] = d.m.Kd[1] = d.m.Kd[2] = 0.5f;
            for (auto &c : n.children) if (c->attr("name") == "reflectance") {
                parse_rgb_map(*c, d.m.Kd, isTex, tex);
                if (isTex) { d.kdTexFile = tex.file; d.sScale = tex.s; d.tScale = tex.t; }
            }
            return d;
        } else if (type == "phong") {
            d.m.type = BSDF_PHONG;
            d.m.Kd[0] = d.m.Kd[1] = d.m.Kd[2] = 0.5f;
            d.m.Ks[0] = d.m.Ks[1] = d.m.Ks[2] = 0.2f;
            d.m.exponent = 30.0f;
            for (auto &c : n.children) {
                const std::string name = c->attr("name");
                if (name == "diffuseReflectance") {
                    parse_rgb_map(*c, d.m.Kd, isTex, tex);
                    if (isTex) { d.kdTexFile = tex.file; d.sScale = tex.s; d.tScale = tex.t; }
                } else if (name == "specularReflectance") {
                    parse_rgb_map(*c, d.m.Ks, isTex, tex);
                    if (isTex) { d.ks.file = tex.file; d.ks.s = tex.s; d.ks.t = tex.t; }
                } else if (name == "exponent") {
                    parse_1d_map(*c, d.m.exponent, d.exponent);
                }
            }
            // GetKsWeight (src/phong.cpp:159-169): texture AVERAGES of Ks and Kd
            const V3 kdAvgV = d.kdTexFile.empty() ? ld3(d.m.Kd) : texture_avg(d.kdTexFile);
            const V3 ksAvgV = d.ks.file.empty() ? ld3(d.m.Ks) : texture_avg(d.ks.file);
            const float ksAvg = luminance(ksAvgV), kdAvg = luminance(kdAvgV);
            const float sum = ksAvg + kdAvg;
            d.m.KsWeight = (sum > 0.0f) ? ksAvg / sum : 0.0f;
            return d;
        } else if (type == "roughdielectric") {
            d.m.type = BSDF_ROUGHDIELECTRIC;
            for (int i = 0; i < 3; i++) { d.m.Ks[i] = 1.0f; d.m.Kt[i] = 1.0f; }
            float intIOR = 1.5046f, extIOR = 1.000277f;
            d.m.alpha = 0.1f;
            for (auto &c : n.children) {
                const std::string name = c->attr("name");
                if (name == "intIOR") intIOR = std::stof(c->attr("value"));
                else if (name == "extIOR") extIOR = std::stof(c->attr("value"));
                else if (name == "alpha") {
                    parse_1d_map(*c, d.m.alpha, d.alpha);
                } else if (name == "specularReflectance") {
                    parse_rgb_map(*c, d.m.Ks, isTex, tex);
                    if (isTex) { d.ks.file = tex.file; d.ks.s = tex.s; d.ks.t = tex.t; }
                } else if (name == "specularTransmittance") {
                    parse_rgb_map(*c, d.m.Kt, isTex, tex);
                    if (isTex) { d.kt.file = tex.file; d.kt.s = tex.s; d.kt.t = tex.t; }
                }
            }
            d.m.eta = intIOR / extIOR;           // src/roughdielectric.h ctor
            d.m.invEta = 1.0f / d.m.eta;
            return d;
        } else if (type == "twosided") {
            for (auto &c : n.children) if (c->name == "bsdf") return parse_bsdf(*c, true);
        }
        throw std::runtime_error("Unknown BSDF: " + type);
    }
};

void load_scene_xml(const std::string &xmlPath, SceneStore &out) {
    std::ifstream f(xmlPath);
    if (!f) throw std::runtime_error("cannot open scene file " + xmlPath);
    std::stringstream buf; buf << f.rdbuf();
    const std::string text = buf.str();
    XmlParser parser(text);
    auto doc = parser.parse();
    const XmlNode *sceneNode = nullptr;
    for (auto &c : doc->children) if (c->name == "scene") sceneNode = c.get();
    if (!sceneNode) throw std::runtime_error("Parse error: no <scene> element");

    out = SceneStore();
    Loader L(out);
    const size_t slash = xmlPath.find_last_of('/');
    L.baseDir = (slash == std::string::npos) ? std::string("") : xmlPath.substr(0, slash + 1);

    Options opt = default_options();
    out.spp = 256; out.directSpp = 256; out.numInitSamples = 300000; out.integrator = "mcmc";
    out.outputName = "image.exr";

    // camera description
    M44 camToWorldM = ident();
    float nearClip = 1e-2f, farClip = 1000.0f, fov = 45.0f;
    int filmW = 512, filmH = 512;

    struct ShapeRec { Mesh mesh; MatDesc mat; bool isLight; V3 radiance; };
    std::vector<ShapeRec> shapes;
    struct LightRec { int type; int shape; V3 a, b; M44 toWorld; std::string file; };
    std::vector<LightRec> lightRecs;

    for (auto &cp : sceneNode->children) {
        const XmlNode &c = *cp;
        if (c.name == "sensor") {
            for (auto &gp : c.children) {
                const XmlNode &g = *gp;
                const std::string name = g.attr("name");
                if (name == "nearClip") nearClip = std::stof(g.attr("value"));
                else if (name == "farClip") farClip = std::stof(g.attr("value"));
                else if (name == "fov") fov = std::stof(g.attr("value"));
                else if (name == "toWorld") {
                    if (g.name == "transform") camToWorldM = parse_transform(g);
                    else throw std::runtime_error("animated camera transforms are not supported (static scenes only)");
                } else if (g.name == "film") {
                    for (auto &hp : g.children) {
                        const std::string hn = hp->attr("name");
                        if (hn == "width") filmW = atoi(hp->attr("value").c_str());
                        else if (hn == "height") filmH = atoi(hp->attr("value").c_str());
                        else if (hn == "filename") out.outputName = hp->attr("value");
                    }
                }
            }
        } else if (c.name == "bsdf") {
            L.bsdfMap[c.attr("id")] = L.parse_bsdf(c, false);
        } else if (c.name == "texture") {
            L.textureMap[c.attr("id")] = L.parse_texture(c);
        } else if (c.name == "emitter") {
            const std::string type = c.attr("type");
            LightRec lr; lr.shape = -1; lr.toWorld = ident();
            if (type == "point") {
                lr.type = LIGHT_POINT; lr.a = mk3s(0.0f); lr.b = mk3s(1.0f);
                for (auto &gp : c.children) {
                    const std::string name = gp->attr("name");
                    if (name == "position") lr.a = mk3(attr_f(*gp, "x", 0), attr_f(*gp, "y", 0), attr_f(*gp, "z", 0));
                    else if (name == "intensity") lr.b = parse_vector3(gp->attr("value"));
                }
            } else if (type == "envmap") {
                lr.type = LIGHT_ENV;
                for (auto &gp : c.children) {
                    const std::string name = gp->attr("name");
                    if (name == "filename") lr.file = gp->attr("value");
                    else if (name == "toWorld") {
                        if (gp->name == "transform") lr.toWorld = parse_transform(*gp);
                        else throw std::runtime_error("animated envmap transforms are not supported");
                    }
                }
            } else throw std::runtime_error("Unsupported emitter");
            lightRecs.push_back(lr);
        } else if (c.name == "shape") {
            ShapeRec sr; sr.isLight = false; sr.radiance = mk3s(1.0f);
            bool haveBsdf = false;
            for (auto &gp : c.children) {
                if (gp->name == "bsdf") { sr.mat = L.parse_bsdf(*gp, false); haveBsdf = true; break; }
                if (gp->name == "ref") {
                    auto it = L.bsdfMap.find(gp->attr("id"));
                    if (it == L.bsdfMap.end()) throw std::runtime_error("ref not found: " + gp->attr("id"));
                    sr.mat = it->second; haveBsdf = true; break;
                }
            }
            if (!haveBsdf) throw std::runtime_error("shape without bsdf");
            std::string filename; int shapeIndex = 0; M44 toWorld = ident();
            bool flipNormals = false, faceNormals = false;
            for (auto &gp : c.children) {
                const std::string name = gp->attr("name");
                if (name == "filename") filename = gp->attr("value");
                else if (name == "shapeIndex") shapeIndex = atoi(gp->attr("value").c_str());
                // sic: the reference converts the attribute's char* to bool, so ANY value
                // (even "false") enables the flag (src/parsescene.cpp:262-265,299-302)
                else if (name == "flipNormals") flipNormals = true;
                else if (name == "faceNormals") faceNormals = true;
                else if (name == "toWorld") {
                    if (gp->name == "transform") toWorld = parse_transform(*gp);
                    else throw std::runtime_error("animated shapes are not supported (static scenes only)");
                }
            }
            const std::string type = c.attr("type");
            if (type == "serialized") sr.mesh = load_serialized(L.baseDir + filename, shapeIndex, toWorld, flipNormals, faceNormals);
            else if (type == "obj") sr.mesh = parse_obj(L.baseDir + filename, toWorld, flipNormals, faceNormals);
            else throw std::runtime_error("Invalid shape type " + type);
            for (auto &gp : c.children) if (gp->name == "emitter") {
                sr.isLight = true;
                for (auto &hp : gp->children) if (hp->attr("name") == "radiance") sr.radiance = parse_vector3(hp->attr("value"));
            }
            shapes.push_back(std::move(sr));
            if (shapes.back().isLight) {
                LightRec lr; lr.type = LIGHT_AREA; lr.shape = (int)shapes.size() - 1; lr.b = shapes.back().radiance; lr.toWorld = ident();
                lightRecs.push_back(lr);
            }
        } else if (c.name == "dpt") {
            opt = default_options();
            for (auto &gp : c.children) {
                const std::string name = gp->attr("name"), v = gp->attr("value");
                if (name == "integrator") out.integrator = v;
                else if (name == "spp") out.spp = std::stoi(v);
                else if (name == "bidirectional") opt.bidirectional = (v == "true");
                else if (name == "numinitsamples") out.numInitSamples = std::stoi(v);
                else if (name == "largestepprob") opt.largeStepProbability = std::stof(v);
                else if (name == "largestepscale") opt.largeStepProbScale = std::stof(v);
                else if (name == "mindepth") opt.minDepth = std::stoi(v);
                else if (name == "maxdepth") opt.maxDepth = std::stoi(v);
                else if (name == "directspp") out.directSpp = std::stoi(v);
                else if (name == "perturbstddev") opt.perturbStdDev = std::stof(v);
                else if (name == "roughnessthreshold") opt.roughnessThreshold = std::stof(v);
                else if (name == "uniformmixprob") opt.uniformMixingProbability = std::stof(v);
                else if (name == "numchains") opt.numChains = std::stoi(v);
                else if (name == "seedoffset") opt.seedOffset = std::stoi(v);
                else if (name == "reportintervalspp") out.reportIntervalSpp = std::stoi(v);
                else if (name == "uselightcoordinatesampling") {
                    // the path sampler here has no light-coordinate branch (src/path.cpp:762-798 / 2974-3013): accepting
                    // the flag would differentiate a different function than the one sampled
                    if (v == "true") throw std::runtime_error("uselightcoordinatesampling is not supported");
                }
                else if (name == "largestepmultiplexed") opt.largeStepMultiplexed = (v == "true");
                else if (name == "h2mc") opt.h2mc = (v == "true");
                else if (name == "mala") opt.mala = (v == "true");
                else if (name == "mala-stepsize") opt.malaStepsize = std::stof(v);
                else if (name == "mala-gn") opt.malaGN = std::stof(v);
                else if (name == "samplecache") { if (v == "true") throw std::runtime_error("samplecache (global cache) is out of scope"); }
                else fprintf(stderr, "Unknown dpt option:%s\n", name.c_str());
            }
        }
    }

    // ---- flatten geometry ----
    struct SrcTri { int geom, prim; };
    std::vector<SrcTri> src;
    Builder B;
    BBoxF sceneBox;
    out.mats.resize(shapes.size());
    for (size_t g = 0; g < shapes.size(); g++) {
        const Mesh &m = shapes[g].mesh;
        MatDesc &md = shapes[g].mat;
        Material mat = md.m;
        if (!md.kdTexFile.empty()) { Loader::TexDesc t; t.file = md.kdTexFile; t.s = md.sScale; t.t = md.tScale; mat.kdTex = L.texture_id(t); }
        auto texId = [&L](const MatTex &mt) { if (mt.file.empty()) return -1; Loader::TexDesc t; t.file = mt.file; t.s = mt.s; t.t = mt.t; return L.texture_id(t); };
        mat.ksTex = texId(md.ks); mat.ktTex = texId(md.kt); mat.expTex = texId(md.exponent); mat.alphaTex = texId(md.alpha);
        mat.hasST = m.st.empty() ? 0 : 1;
        mat.areaLight = -1; mat.invTotalArea = 0.0f; mat.firstTid = 0;
        out.mats[g] = mat;
        for (auto &p : m.pos) { const float q[3] = {p.x, p.y, p.z}; sceneBox.grow(q); }
        for (size_t t = 0; t < m.idx.size() / 3; t++) {
            BuildPrim bp;
            for (int k = 0; k < 3; k++) { const V3 &p = m.pos[m.idx[3 * t + k]]; const float q[3] = {p.x, p.y, p.z}; bp.box.grow(q); }
            for (int a = 0; a < 3; a++) bp.c[a] = 0.5f * (bp.box.mn[a] + bp.box.mx[a]);
            bp.src = (int)src.size();
            src.push_back({(int)g, (int)t});
            B.prims.push_back(bp);
        }
    }
    if (src.empty()) throw std::runtime_error("scene has no triangles");
    BBoxF rootBox;
    const int root = B.build(0, (int)B.prims.size(), rootBox, 0);
    if (B.maxDepth > LMC_BVH_STACK - 1)
        throw std::runtime_error("BVH deeper than the traversal stack (" + std::to_string(B.maxDepth) + " levels)");
    if (root < 0) {
        // single leaf: wrap it in a node so traversal always starts at an inner node
        BvhNode nd; memset(&nd, 0, sizeof(nd));
        for (int i = 0; i < 3; i++) { nd.lmin[i] = rootBox.mn[i]; nd.lmax[i] = rootBox.mx[i]; nd.rmin[i] = INFINITY; nd.rmax[i] = -INFINITY; }
        nd.left = root; nd.right = root;
        // an empty right box never passes box_test
        B.nodes.push_back(nd);
    }
    // Breadth-first node order: the first K nodes are the top levels of the tree (the traversal kernels
    // stage them in shared memory); child indices are remapped, triangle order (= tid) is untouched.
    {
        std::vector<int> bfs; bfs.reserve(B.nodes.size());
        std::vector<int> newIdx(B.nodes.size(), -1);
        bfs.push_back(root < 0 ? 0 : root);
        for (size_t head = 0; head < bfs.size(); head++) {
            const BvhNode &nd = B.nodes[bfs[head]];
            newIdx[bfs[head]] = (int)head;
            if (nd.left >= 0) bfs.push_back(nd.left);
            if (nd.right >= 0 && nd.right != nd.left) bfs.push_back(nd.right);
        }
        std::vector<BvhNode> re(bfs.size());
        for (size_t k = 0; k < bfs.size(); k++) {
            BvhNode nd = B.nodes[bfs[k]];
            if (nd.left >= 0) nd.left = newIdx[nd.left];
            if (nd.right >= 0) nd.right = newIdx[nd.right];
            re[k] = nd;
        }
        B.nodes.swap(re);
    }
    out.nodes = B.nodes;
    const int nt = (int)B.order.size();
    out.tris.resize(nt); out.shade.resize(nt);
    std::vector<std::vector<int>> primToTid(shapes.size());
    for (size_t g = 0; g < shapes.size(); g++) primToTid[g].assign(shapes[g].mesh.idx.size() / 3, -1);
    for (int tid = 0; tid < nt; tid++) {
        const SrcTri s = src[B.order[tid]];
        const Mesh &m = shapes[s.geom].mesh;
        const uint32_t i0 = m.idx[3 * s.prim], i1 = m.idx[3 * s.prim + 1], i2 = m.idx[3 * s.prim + 2];
        const V3 p0 = m.pos[i0], e1 = m.pos[i1] - m.pos[i0], e2 = m.pos[i2] - m.pos[i0];
        TriGeom &tg = out.tris[tid];
        tg.p0[0] = p0.x; tg.p0[1] = p0.y; tg.p0[2] = p0.z; tg.geom = s.geom;
        tg.e1[0] = e1.x; tg.e1[1] = e1.y; tg.e1[2] = e1.z; tg.prim = s.prim;
        tg.e2[0] = e2.x; tg.e2[1] = e2.y; tg.e2[2] = e2.z; tg.pad = 0;
        TriShade &ts = out.shade[tid];
        memset(&ts, 0, sizeof(ts));
        const V3 n0 = m.nor[i0], n1 = m.nor[i1], n2 = m.nor[i2];
        ts.n0[0] = n0.x; ts.n0[1] = n0.y; ts.n0[2] = n0.z;
        ts.n1[0] = n1.x; ts.n1[1] = n1.y; ts.n1[2] = n1.z;
        ts.n2[0] = n2.x; ts.n2[1] = n2.y; ts.n2[2] = n2.z;
        if (!m.st.empty()) {
            ts.st0[0] = m.st[i0].x; ts.st0[1] = m.st[i0].y; ts.st1[0] = m.st[i1].x; ts.st1[1] = m.st[i1].y;
            ts.st2[0] = m.st[i2].x; ts.st2[1] = m.st[i2].y;
        }
        primToTid[s.geom][s.prim] = tid;
    }

    // ---- lights (XML order) ----
    std::vector<float> weights;
    out.head.env.present = 0; out.head.env.lightIndex = -1;
    for (size_t li = 0; li < lightRecs.size(); li++) {
        const LightRec &lr = lightRecs[li];
        Light l; memset(&l, 0, sizeof(l));
        l.type = lr.type; l.samplingWeight = 1.0f; l.geom = -1;
        if (lr.type == LIGHT_POINT) {
            l.pos[0] = lr.a.x; l.pos[1] = lr.a.y; l.pos[2] = lr.a.z;
            l.emission[0] = lr.b.x; l.emission[1] = lr.b.y; l.emission[2] = lr.b.z;
        } else if (lr.type == LIGHT_AREA) {
            const Mesh &m = shapes[lr.shape].mesh;
            const int np = (int)m.idx.size() / 3;
            std::vector<float> area(np);
            float totalArea = 0.0f;
            for (int i = 0; i < np; i++) {
                const V3 p0 = m.pos[m.idx[3 * i]], p1 = m.pos[m.idx[3 * i + 1]], p2 = m.pos[m.idx[3 * i + 2]];
                area[i] = 0.5f * length(cross(p1 - p0, p2 - p0));
                totalArea += area[i];
            }
            l.geom = lr.shape; l.numPrims = np;
            l.primCdfOffset = (int)out.lightCdf.size();
            const std::vector<float> cdf = piecewise_cdf(area);
            out.lightCdf.insert(out.lightCdf.end(), cdf.begin(), cdf.end());
            l.primTidOffset = (int)out.lightPrimTid.size();
            out.lightPrimTid.insert(out.lightPrimTid.end(), primToTid[lr.shape].begin(), primToTid[lr.shape].end());
            l.emission[0] = lr.b.x; l.emission[1] = lr.b.y; l.emission[2] = lr.b.z;
            l.invTotalArea = 1.0f / totalArea;
            out.mats[lr.shape].areaLight = (int)li;
            out.mats[lr.shape].invTotalArea = l.invTotalArea;
        } else {
            if (out.head.env.present) throw std::runtime_error("more than one envmap");
            const RawImage &im = L.image(lr.file);
            EnvMap &e = out.head.env;
            e.present = 1; e.lightIndex = (int)li; e.width = im.w; e.height = im.h;
            out.envImage = im.rgb;
            const int W = im.w, H = im.h;
            out.envCdfCols.assign((size_t)(W + 1) * H, 0.0f);
            out.envCdfRows.assign(H + 1, 0.0f);
            out.envRowWeights.assign(H, 0.0f);
            size_t colPos = 0, rowPos = 0;
            float rowSum = 0.0f;
            out.envCdfRows[rowPos++] = 0.0f;
            for (int y = 0; y < H; y++) {
                float colSum = 0.0f;
                out.envCdfCols[colPos++] = 0.0f;
                for (int x = 0; x < W; x++) {
                    colSum += luminance(ld3(&im.rgb[3 * ((size_t)y * W + x)]));
                    out.envCdfCols[colPos++] = colSum;
                }
                const float normalization = 1.0f / colSum;
                for (int x = 1; x < W; x++) out.envCdfCols[colPos - x - 1] *= normalization;
                out.envCdfCols[colPos - 1] = 1.0f;
                const float weight = sinf(((float)y + 0.5f) * LMC_PI / (float)H);
                out.envRowWeights[y] = weight;
                rowSum += colSum * weight;
                out.envCdfRows[rowPos++] = rowSum;
            }
            float normalization = 1.0f / rowSum;
            for (int y = 1; y < H; y++) out.envCdfRows[rowPos - y - 1] *= normalization;
            out.envCdfRows[rowPos - 1] = 1.0f;
            if (rowSum == 0 || !std::isfinite(rowSum)) throw std::runtime_error("Invalid environment map");
            e.normalization = 1.0f / (rowSum * (LMC_TWOPI / (float)W) * (LMC_PI / (float)H));
            e.pixelSize[0] = LMC_TWOPI / (float)W;
            e.pixelSize[1] = (float)(M_PI / (double)H);
            const Rigid r = decompose(lr.toWorld);
            const Rigid ri = invert_rigid(r);
            e.toWorld = rigid_to_m(r); e.toLight = rigid_to_m(ri);
            rigid_serialize(r, e.toWorldSer); rigid_serialize(ri, e.toLightSer);
        }
        weights.push_back(l.samplingWeight);
        out.lights.push_back(l);
    }
    if (out.lights.empty()) throw std::runtime_error("scene has no lights");
    out.lightPickCdf = piecewise_cdf(weights);
    out.head.lightWeightSum = 0.0f;
    for (float w : weights) out.head.lightWeightSum += w;

    // ---- camera ----
    Camera &cam = out.head.cam;
    const Rigid cr = decompose(camToWorldM);
    const Rigid cri = invert_rigid(cr);
    cam.camToWorld = rigid_to_m(cr); cam.worldToCam = rigid_to_m(cri);
    rigid_serialize(cr, cam.camToWorldSer);
    const float aspect = (float)filmW / (float)filmH;
    cam.camToSample = mul(mul(scale_m(-0.5f, -0.5f * aspect, 1.0f), translate_m(-1.0f, -1.0f / aspect, 0.0f)),
                          perspective_m(fov, nearClip, farClip));
    cam.sampleToCam = inverse44(cam.camToSample);
    cam.dist = (float)filmW / (2.0f * tanf((fov / 2.0f) * (LMC_PI / 180.0f)));
    cam.nearClip = nearClip; cam.farClip = farClip; cam.width = filmW; cam.height = filmH;

    // ---- bounding sphere (x1000) ----
    const V3 mn = mk3(sceneBox.mn[0], sceneBox.mn[1], sceneBox.mn[2]), mx = mk3(sceneBox.mx[0], sceneBox.mx[1], sceneBox.mx[2]);
    const V3 center = 0.5f * (mn + mx);
    out.head.bsphereCenter[0] = center.x; out.head.bsphereCenter[1] = center.y; out.head.bsphereCenter[2] = center.z;
    out.head.bsphereRadius = 0.5f * dm_sqrt(distance_squared(mn, mx)) * 1000.0f;

    // ---- Serialize(scene) (src/scene.cpp:164-169; matrices column-major, src/utils.h:347-354) ----
    float *s = out.head.sceneSer; int k = 0;
    s[k++] = opt.useLightCoordinateSampling ? 1.0f : 0.0f;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) s[k++] = cam.sampleToCam.m[r][c];
    for (int i = 0; i < 15; i++) s[k++] = cam.camToWorldSer[i];
    s[k++] = (float)(filmH * filmW);
    s[k++] = cam.dist;
    s[k++] = center.x; s[k++] = center.y; s[k++] = center.z; s[k++] = out.head.bsphereRadius;

    out.head.opt = opt;
    out.head.numTris = nt; out.head.numNodes = (int)out.nodes.size(); out.head.numGeoms = (int)out.mats.size();
    out.head.numTextures = (int)out.textures.size(); out.head.numLights = (int)out.lights.size();
}

Scene SceneStore::view() const {
    Scene s = head;
    s.tris = tris.data(); s.shade = shade.data(); s.nodes = nodes.data(); s.mats = mats.data();
    s.textures = textures.data(); s.texData = texData.data();
    s.lights = lights.data(); s.lightPickCdf = lightPickCdf.data();
    s.lightCdf = lightCdf.data(); s.lightPrimTid = lightPrimTid.data();
    s.env.image = envImage.data(); s.env.cdfRows = envCdfRows.data(); s.env.cdfCols = envCdfCols.data();
    s.env.rowWeights = envRowWeights.data();
    return s;
}

// ------------------------------------------------------------------------------------------
// scene pack: [magic 'LMCP' u32 version] head(Scene, pointers zeroed) + length-prefixed arrays
// ------------------------------------------------------------------------------------------
template <class T> static void wr_vec(std::ofstream &f, const std::vector<T> &v) {
    const uint64_t n = v.size(); f.write((const char *)&n, 8);
    if (n) f.write((const char *)v.data(), n * sizeof(T));
}
template <class T> static void rd_vec(std::ifstream &f, std::vector<T> &v) {
    uint64_t n = 0; f.read((char *)&n, 8);
    if (!f || n > (1ull << 32)) throw std::runtime_error("corrupt scene pack");
    v.resize(n);
    if (n) f.read((char *)v.data(), n * sizeof(T));
}
static const uint32_t kPackVersion = 3;

void save_scene_pack(const std::string &path, const SceneStore &s) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    f.write("LMCP", 4); f.write((const char *)&kPackVersion, 4);
    const uint32_t headSize = sizeof(Scene); f.write((const char *)&headSize, 4);
    Scene h = s.head;
    h.tris = nullptr; h.shade = nullptr; h.nodes = nullptr; h.mats = nullptr; h.textures = nullptr; h.texData = nullptr;
    h.lights = nullptr; h.lightPickCdf = nullptr; h.lightCdf = nullptr; h.lightPrimTid = nullptr;
    h.env.image = nullptr; h.env.cdfRows = nullptr; h.env.cdfCols = nullptr; h.env.rowWeights = nullptr;
    f.write((const char *)&h, sizeof(h));
    wr_vec(f, s.tris); wr_vec(f, s.shade); wr_vec(f, s.nodes); wr_vec(f, s.mats); wr_vec(f, s.textures);
    wr_vec(f, s.texData); wr_vec(f, s.lights); wr_vec(f, s.lightPickCdf); wr_vec(f, s.lightCdf);
    wr_vec(f, s.lightPrimTid); wr_vec(f, s.envImage); wr_vec(f, s.envCdfRows); wr_vec(f, s.envCdfCols);
    wr_vec(f, s.envRowWeights);
    const int meta[3] = {s.spp, s.directSpp, s.numInitSamples}; f.write((const char *)meta, 12);
}

// number of inner nodes on the longest root-to-leaf path (children < 0 are leaves); -1 for a malformed array
int bvh_depth(const std::vector<BvhNode> &nodes) {
    if (nodes.empty()) return 0;
    std::vector<std::pair<int, int>> st;
    st.push_back({0, 1});
    int deepest = 0; size_t visited = 0;
    while (!st.empty()) {
        const std::pair<int, int> t = st.back(); st.pop_back();
        if (t.first < 0 || t.first >= (int)nodes.size() || ++visited > 2 * nodes.size()) return -1;
        if (t.second > deepest) deepest = t.second;
        const BvhNode &nd = nodes[t.first];
        if (nd.left >= 0) st.push_back({nd.left, t.second + 1});
        if (nd.right >= 0 && nd.right != nd.left) st.push_back({nd.right, t.second + 1});
    }
    return deepest;
}

void load_scene_pack(const std::string &path, SceneStore &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open scene pack " + path);
    char magic[4]; uint32_t ver = 0, headSize = 0;
    f.read(magic, 4); f.read((char *)&ver, 4); f.read((char *)&headSize, 4);
    if (memcmp(magic, "LMCP", 4) != 0 || ver != kPackVersion || headSize != sizeof(Scene))
        throw std::runtime_error("scene pack version/layout mismatch: " + path);
    out = SceneStore();
    f.read((char *)&out.head, sizeof(Scene));
    rd_vec(f, out.tris); rd_vec(f, out.shade); rd_vec(f, out.nodes); rd_vec(f, out.mats); rd_vec(f, out.textures);
    rd_vec(f, out.texData); rd_vec(f, out.lights); rd_vec(f, out.lightPickCdf); rd_vec(f, out.lightCdf);
    rd_vec(f, out.lightPrimTid); rd_vec(f, out.envImage); rd_vec(f, out.envCdfRows); rd_vec(f, out.envCdfCols);
    rd_vec(f, out.envRowWeights);
    int meta[3]; f.read((char *)meta, 12);
    if (!f) throw std::runtime_error("truncated scene pack " + path);
    const int depth = bvh_depth(out.nodes);
    if (depth < 0 || depth > LMC_BVH_STACK)
        throw std::runtime_error("scene pack: BVH malformed or deeper than the traversal stack: " + path);
    out.spp = meta[0]; out.directSpp = meta[1]; out.numInitSamples = meta[2];
    out.integrator = "mcmc";
}

}  // namespace lmc_host
