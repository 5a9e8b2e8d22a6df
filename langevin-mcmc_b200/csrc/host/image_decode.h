// image_decode.h -- in-loader decoding of the image formats the reference's scenes use: PNG and JPEG textures
// (JPEG: jpeg_decode.h) and OpenEXR environment maps.
//
// The reference reads images through OpenImageIO (src/image.cpp:5-45, src/bitmaptexture.h:73-146), which is vendored
// only as unbuilt source.  These two decoders follow the published file formats directly (PNG: RFC 2083; OpenEXR:
// "OpenEXR File Layout") on top of zlib's inflate, which the loader links anyway for .serialized meshes:
//   PNG   8-bit grey / grey+alpha / RGB / RGBA / palette, non-interlaced
//   EXR   single-part scan-line files, compression NONE / ZIPS / ZIP, HALF or FLOAT channels R G B (or Y)
// Both are lossless, so the result is bit-identical to the OpenCV decode kept in tests/golden/decoded/*.rawf
// (tests/test_loader_bvh.py); the JPEG decoder reproduces the libjpeg family's default arithmetic and matches them too.
#pragma once
#include <zlib.h>
#include <stdint.h>
#include <string.h>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "jpeg_decode.h"

namespace lmc_host {

struct DecodedImage { int w = 0, h = 0, is8 = 0; std::vector<float> rgb; };   // row 0 = top, RGB interleaved

namespace imgdetail {
inline std::vector<unsigned char> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open image " + path);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline std::vector<unsigned char> inflate_all(const unsigned char *src, size_t n, size_t expect) {
    std::vector<unsigned char> out(expect);
    uLongf len = (uLongf)expect;
    const int rc = uncompress(out.data(), &len, src, (uLong)n);
    if (rc != Z_OK) throw std::runtime_error("image: zlib inflate failed");
    out.resize(len);
    return out;
}
inline uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = s;
        else { int k = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; k++; } u = s | ((uint32_t)(113 - k) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4);
    return f;
}
}  // namespace imgdetail

inline DecodedImage decode_png(const std::string &path) {
    using namespace imgdetail;
    const std::vector<unsigned char> b = read_file(path);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (b.size() < 33 || memcmp(b.data(), sig, 8) != 0) throw std::runtime_error("not a PNG file: " + path);
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte;
    for (size_t o = 8; o + 12 <= b.size();) {
        const uint32_t len = be32(&b[o]);
        const char *type = (const char *)&b[o + 4];
        const unsigned char *d = &b[o + 8];
        if (o + 12 + len > b.size()) throw std::runtime_error("truncated PNG: " + path);
        if (!memcmp(type, "IHDR", 4)) { w = (int)be32(d); h = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        o += 12 + len;
    }
    if (depth != 8 || interlace != 0) throw std::runtime_error("PNG: only 8-bit non-interlaced images are supported: " + path);
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw std::runtime_error("PNG: unknown colour type: " + path);
    const size_t stride = (size_t)w * ch;
    std::vector<unsigned char> raw = inflate_all(idat.data(), idat.size(), (stride + 1) * (size_t)h);
    if (raw.size() != (stride + 1) * (size_t)h) throw std::runtime_error("PNG: unexpected data size: " + path);
    std::vector<unsigned char> pix(stride * h);
    for (int y = 0; y < h; y++) {
        const unsigned char *src = &raw[(stride + 1) * y];
        unsigned char *dst = &pix[stride * y];
        const unsigned char *up = y ? &pix[stride * (y - 1)] : nullptr;
        const int ft = src[0];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)ch ? dst[x - ch] : 0, bb = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = bb;
            else if (ft == 3) pred = (a + bb) >> 1;
            else if (ft == 4) { const int p = a + bb - c, pa = abs(p - a), pb = abs(p - bb), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c); }
            else if (ft != 0) throw std::runtime_error("PNG: bad filter type: " + path);
            dst[x] = (unsigned char)(src[1 + x] + pred);
        }
    }
    DecodedImage im; im.w = w; im.h = h; im.is8 = 1; im.rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        unsigned char r, g, bl;
        if (ctype == 0 || ctype == 4) r = g = bl = pix[i * ch];
        else if (ctype == 3) { const size_t k = 3 * (size_t)pix[i]; if (k + 2 >= plte.size()) throw std::runtime_error("PNG: palette index out of range"); r = plte[k]; g = plte[k + 1]; bl = plte[k + 2]; }
        else { r = pix[i * ch]; g = pix[i * ch + 1]; bl = pix[i * ch + 2]; }
        im.rgb[3 * i] = (float)r / 255.0f; im.rgb[3 * i + 1] = (float)g / 255.0f; im.rgb[3 * i + 2] = (float)bl / 255.0f;
    }
    return im;
}

inline DecodedImage decode_exr(const std::string &path) {
    using namespace imgdetail;
    const std::vector<unsigned char> b = read_file(path);
    if (b.size() < 16 || b[0] != 0x76 || b[1] != 0x2f || b[2] != 0x31 || b[3] != 0x01) throw std::runtime_error("not an OpenEXR file: " + path);
    if (b[5] & 0x1a) throw std::runtime_error("EXR: tiled / multi-part / deep files are not supported: " + path);
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    int comp = -1, x0 = 0, y0 = 0, x1 = -1, y1 = -1, lineOrder = 0;
    size_t o = 8;
    while (o < b.size() && b[o] != 0) {
        const std::string name((const char *)&b[o]); o += name.size() + 1;
        const std::string type((const char *)&b[o]); o += type.size() + 1;
        int32_t size; memcpy(&size, &b[o], 4); o += 4;
        const unsigned char *v = &b[o];
        if (name == "channels") {
            for (size_t p = 0; v[p] != 0;) {
                Chan c; c.name = (const char *)&v[p]; p += c.name.size() + 1;
                int32_t t; memcpy(&t, &v[p], 4); c.type = t; p += 16;
                chans.push_back(c);
            }
        } else if (name == "compression") comp = v[0];
        else if (name == "dataWindow") { int32_t w4[4]; memcpy(w4, v, 16); x0 = w4[0]; y0 = w4[1]; x1 = w4[2]; y1 = w4[3]; }
        else if (name == "lineOrder") lineOrder = v[0];
        o += size;
    }
    o += 1;
    if (comp != 0 && comp != 2 && comp != 3) throw std::runtime_error("EXR: only NONE / ZIPS / ZIP compression is supported: " + path);
    const int w = x1 - x0 + 1, h = y1 - y0 + 1;
    if (w <= 0 || h <= 0 || chans.empty()) throw std::runtime_error("EXR: bad header: " + path);
    (void)lineOrder;   // blocks carry their y coordinate; the offset table is walked in file order
    int ci[3] = {-1, -1, -1};
    size_t lineBytes = 0;
    std::vector<size_t> chanOff(chans.size());
    for (size_t c = 0; c < chans.size(); c++) {
        if (chans[c].type != 1 && chans[c].type != 2) throw std::runtime_error("EXR: only HALF / FLOAT channels are supported: " + path);
        chanOff[c] = lineBytes;
        lineBytes += (size_t)w * (chans[c].type == 1 ? 2 : 4);
        if (chans[c].name == "R") ci[0] = (int)c; else if (chans[c].name == "G") ci[1] = (int)c; else if (chans[c].name == "B") ci[2] = (int)c;
        else if (chans[c].name == "Y") ci[0] = ci[1] = ci[2] = (int)c;
    }
    if (ci[0] < 0 || ci[1] < 0 || ci[2] < 0) throw std::runtime_error("EXR: needs R, G, B (or Y) channels: " + path);
    const int linesPerBlock = comp == 3 ? 16 : 1;
    const int nBlocks = (h + linesPerBlock - 1) / linesPerBlock;
    DecodedImage im; im.w = w; im.h = h; im.is8 = 0; im.rgb.assign((size_t)w * h * 3, 0.0f);
    for (int blk = 0; blk < nBlocks; blk++) {
        uint64_t off; memcpy(&off, &b[o + 8 * (size_t)blk], 8);
        if (off + 8 > b.size()) throw std::runtime_error("truncated EXR: " + path);
        int32_t y, size; memcpy(&y, &b[off], 4); memcpy(&size, &b[off + 4], 4);
        const int lines = std::min(linesPerBlock, y1 - y + 1);
        const size_t expect = lineBytes * (size_t)lines;
        std::vector<unsigned char> data;
        if (comp == 0 || (size_t)size == expect) data.assign(&b[off + 8], &b[off + 8] + size);
        else {
            std::vector<unsigned char> t = inflate_all(&b[off + 8], (size_t)size, expect);
            if (t.size() != expect) throw std::runtime_error("EXR: unexpected block size: " + path);
            for (size_t i = 1; i < t.size(); i++) t[i] = (unsigned char)(t[i - 1] + t[i] - 128);      // predictor
            data.resize(expect);
            const size_t half = (expect + 1) / 2;                                                      // de-interleave
            for (size_t i = 0; i < expect; i++) data[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
        }
        for (int l = 0; l < lines; l++) {
            const int row = y - y0 + l;
            const unsigned char *line = &data[lineBytes * (size_t)l];
            for (int k = 0; k < 3; k++) {
                const Chan &c = chans[ci[k]];
                const unsigned char *src = line + chanOff[ci[k]];
                for (int x = 0; x < w; x++) {
                    float f;
                    if (c.type == 1) { uint16_t hv; memcpy(&hv, src + 2 * (size_t)x, 2); f = half_to_float(hv); }
                    else memcpy(&f, src + 4 * (size_t)x, 4);
                    im.rgb[((size_t)row * w + x) * 3 + k] = f;
                }
            }
        }
    }
    return im;
}

inline DecodedImage decode_jpeg(const std::string &path) {
    const JpegImage j = decode_jpeg_bytes(imgdetail::read_file(path), path);
    DecodedImage im; im.w = j.w; im.h = j.h; im.is8 = 1; im.rgb.resize(j.rgb.size());
    for (size_t i = 0; i < j.rgb.size(); i++) im.rgb[i] = (float)j.rgb[i] / 255.0f;
    return im;
}

// PNG / EXR / JPEG by extension
inline bool decode_image_native(const std::string &path, DecodedImage &out) {
    const size_t dot = path.rfind('.');
    std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
    for (auto &c : ext) c = (char)tolower(c);
    if (ext == "png") { out = decode_png(path); return true; }
    if (ext == "exr") { out = decode_exr(path); return true; }
    if (ext == "jpg" || ext == "jpeg") { out = decode_jpeg(path); return true; }
    return false;
}

}  // namespace lmc_host
