// image_io.h -- the film side of the render loop after the chains: MergeBuffer / BufferToFilm and WriteImage.
//
// Reference: MergeBuffer + BufferToFilm src/image.h:79-105 (film = b1Weight * buffer1 + b2Weight * buffer2, as
// called at src/mlt.cpp:203-207 with 1/directSpp and 1/spp), WriteImage src/image.cpp:29-60 (OpenImageIO, float
// EXR).  OIIO is not vendored in a buildable form, so the writer below emits the OpenEXR file format directly:
// version 2, single part, scan lines, NO_COMPRESSION, three 32-bit float channels B, G, R -- readable by any EXR
// reader (tests read it back with an independent parser and with OpenCV).  `.pfm` is written for the same data
// when the file name asks for it.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace lmc_host {

inline void merge_buffer(const float *buffer1, float b1Weight, const float *buffer2, float b2Weight, long long n, float *film) {
    for (long long i = 0; i < n; i++) film[i] = b1Weight * (buffer1 ? buffer1[i] : 0.0f) + b2Weight * (buffer2 ? buffer2[i] : 0.0f);
}

namespace exr_detail {
inline void put(std::vector<unsigned char> &b, const void *p, size_t n) { const unsigned char *c = (const unsigned char *)p; b.insert(b.end(), c, c + n); }
inline void put_str(std::vector<unsigned char> &b, const char *s) { put(b, s, strlen(s) + 1); }
inline void put_i32(std::vector<unsigned char> &b, int32_t v) { put(b, &v, 4); }
inline void put_f32(std::vector<unsigned char> &b, float v) { put(b, &v, 4); }
inline void attr(std::vector<unsigned char> &b, const char *name, const char *type, const std::vector<unsigned char> &value) {
    put_str(b, name); put_str(b, type); put_i32(b, (int32_t)value.size()); put(b, value.data(), value.size());
}
}  // namespace exr_detail

// rgb: H x W x 3 floats, row 0 = top scan line.  Returns false on I/O failure.
inline bool write_exr(const std::string &path, int w, int h, const float *rgb) {
    using namespace exr_detail;
    std::vector<unsigned char> hd;
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(hd, magic, 8);
    {   // chlist: name, pixel type FLOAT (2), pLinear, reserved[3], xSampling, ySampling; alphabetical order
        std::vector<unsigned char> v;
        for (const char *ch : {"B", "G", "R"}) {
            put_str(v, ch); put_i32(v, 2);
            const unsigned char lin[4] = {0, 0, 0, 0}; put(v, lin, 4);
            put_i32(v, 1); put_i32(v, 1);
        }
        v.push_back(0);
        attr(hd, "channels", "chlist", v);
    }
    { std::vector<unsigned char> v(1, 0); attr(hd, "compression", "compression", v); }
    for (const char *name : {"dataWindow", "displayWindow"}) {
        std::vector<unsigned char> v; put_i32(v, 0); put_i32(v, 0); put_i32(v, w - 1); put_i32(v, h - 1);
        attr(hd, name, "box2i", v);
    }
    { std::vector<unsigned char> v(1, 0); attr(hd, "lineOrder", "lineOrder", v); }
    { std::vector<unsigned char> v; put_f32(v, 1.0f); attr(hd, "pixelAspectRatio", "float", v); }
    { std::vector<unsigned char> v; put_f32(v, 0.0f); put_f32(v, 0.0f); attr(hd, "screenWindowCenter", "v2f", v); }
    { std::vector<unsigned char> v; put_f32(v, 1.0f); attr(hd, "screenWindowWidth", "float", v); }
    hd.push_back(0);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint64_t lineBytes = 8 + (uint64_t)w * 3 * 4;
    const uint64_t first = hd.size() + (uint64_t)h * 8;
    bool ok = fwrite(hd.data(), 1, hd.size(), f) == hd.size();
    for (int y = 0; y < h && ok; y++) { const uint64_t off = first + (uint64_t)y * lineBytes; ok = fwrite(&off, 8, 1, f) == 1; }
    std::vector<float> line((size_t)w * 3);
    for (int y = 0; y < h && ok; y++) {
        const int32_t head[2] = {y, (int32_t)(w * 3 * 4)};
        const float *row = rgb + (size_t)y * w * 3;
        for (int x = 0; x < w; x++) { line[x] = row[3 * x + 2]; line[w + x] = row[3 * x + 1]; line[2 * w + x] = row[3 * x]; }
        ok = fwrite(head, 4, 2, f) == 2 && fwrite(line.data(), 4, line.size(), f) == line.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok;
}

// Portable float map: "PF\n<w> <h>\n-1.0\n", rows bottom to top, little endian
inline bool write_pfm(const std::string &path, int w, int h, const float *rgb) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
    bool ok = true;
    for (int y = h - 1; y >= 0 && ok; y--) ok = fwrite(rgb + (size_t)y * w * 3, 4, (size_t)w * 3, f) == (size_t)w * 3;
    ok = (fclose(f) == 0) && ok;
    return ok;
}

inline bool write_image(const std::string &path, int w, int h, const float *rgb) {
    if (path.size() > 4 && path.substr(path.size() - 4) == ".pfm") return write_pfm(path, w, h, rgb);
    return write_exr(path, w, h, rgb);
}

}  // namespace lmc_host
