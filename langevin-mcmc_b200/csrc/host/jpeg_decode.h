// jpeg_decode.h -- in-loader JPEG decoder (ITU-T T.81): baseline / extended sequential (SOF0 / SOF1) and progressive
// (SOF2) Huffman files, 8-bit, 1 or 3 components, chroma subsampled 1x1, 2x1 or 2x2.
//
// The reference reads its textures through OpenImageIO (src/bitmaptexture.h:73-146), i.e. through libjpeg.  JPEG is
// lossy and a decoder is free in its IDCT and chroma upsampling, so "the same texture" only exists relative to one
// decoder: this one follows the published algorithms the libjpeg family uses by default -- the 13-bit fixed-point
// Loeffler-Ligtenberg-Moschytz inverse DCT ("islow"), the 3:1 triangle-filter ("fancy") chroma upsampling with edge
// replication at the component's real extent, and the 16-bit fixed-point YCbCr -> RGB tables of JFIF -- restated from
// their specifications, and is checked to return, for every JPEG of the bundled scenes, exactly the bytes OpenCV's
// decoder returns (tests/golden/decoded/*.rawf, tests/test_loader_bvh.py).
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace lmc_host {
namespace jpgdetail {

static const unsigned char kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                          41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                          30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    unsigned char bits[17] = {0};
    unsigned char vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct BitReader {
    const unsigned char *p, *end;
    uint32_t acc = 0; int n = 0;
    bool hitMarker = false;
    BitReader(const unsigned char *b, const unsigned char *e) : p(b), end(e) {}
    void fill() {
        while (n <= 24) {
            int c = 0;
            if (!hitMarker && p < end) {
                c = *p;
                if (c == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { hitMarker = true; c = 0; }       // a marker ends the entropy-coded segment: feed zeros
                } else p++;
            }
            acc |= (uint32_t)c << (24 - n);
            n += 8;
        }
    }
    int get(int k) {
        if (k == 0) return 0;
        if (n < k) fill();
        const int v = (int)(acc >> (32 - k));
        acc <<= k; n -= k;
        return v;
    }
    int decode(const Huff &h) {
        int code = 0;
        for (int l = 1; l <= 16; l++) {
            code = (code << 1) | get(1);
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        throw std::runtime_error("JPEG: bad Huffman code");
    }
    // byte-align and step over an RSTn marker
    void restart() {
        acc = 0; n = 0;
        if (hitMarker) { hitMarker = false; }
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) p++;
        if (p + 1 < end) p += 2;
    }
};

inline int extend(int r, int s) { return r < (1 << (s - 1)) ? r + (int)((~0u) << s) + 1 : r; }

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0;
    int bw = 0, bh = 0;          // blocks per row / rows in the (MCU-padded) coefficient array
    int cw = 0, chh = 0;         // real extent in blocks (non-interleaved scans walk this)
    int dw = 0, dh = 0;          // real extent in samples ("downsampled" width / height)
    std::vector<short> coef;     // bw * bh blocks of 64, natural order
    std::vector<unsigned char> plane;   // (bw * 8) x (bh * 8) samples after the inverse DCT
    int dcTab = 0, acTab = 0, pred = 0;
};

// 13-bit fixed-point LL&M inverse DCT with the libjpeg scaling ("islow": CONST_BITS 13, PASS1_BITS 2)
inline void idct_islow(const short *in, const uint16_t *q, unsigned char *out, int stride) {
    const int64_t F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                  F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
    int ws[64];
    for (int c = 0; c < 8; c++) {
        int64_t z2 = (int)in[16 + c] * (int)q[16 + c], z3 = (int)in[48 + c] * (int)q[48 + c];
        int64_t z1 = (z2 + z3) * F0_541;
        int64_t tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        z2 = (int)in[c] * (int)q[c]; z3 = (int)in[32 + c] * (int)q[32 + c];
        int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
        const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = (int)in[56 + c] * (int)q[56 + c]; tmp1 = (int)in[40 + c] * (int)q[40 + c];
        tmp2 = (int)in[24 + c] * (int)q[24 + c]; tmp3 = (int)in[8 + c] * (int)q[8 + c];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; int64_t z4 = tmp1 + tmp3;
        const int64_t z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int64_t r = 1 << 10;
        ws[c] = (int)((tmp10 + tmp3 + r) >> 11); ws[56 + c] = (int)((tmp10 - tmp3 + r) >> 11);
        ws[8 + c] = (int)((tmp11 + tmp2 + r) >> 11); ws[48 + c] = (int)((tmp11 - tmp2 + r) >> 11);
        ws[16 + c] = (int)((tmp12 + tmp1 + r) >> 11); ws[40 + c] = (int)((tmp12 - tmp1 + r) >> 11);
        ws[24 + c] = (int)((tmp13 + tmp0 + r) >> 11); ws[32 + c] = (int)((tmp13 - tmp0 + r) >> 11);
    }
    for (int rrow = 0; rrow < 8; rrow++) {
        const int *w = ws + 8 * rrow;
        int64_t z2 = w[2], z3 = w[6];
        int64_t z1 = (z2 + z3) * F0_541;
        int64_t tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
        int64_t tmp0 = ((int64_t)w[0] + w[4]) * 8192, tmp1 = ((int64_t)w[0] - w[4]) * 8192;
        const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; int64_t z4 = tmp1 + tmp3;
        const int64_t z5 = (z3 + z4) * F1_175;
        tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
        z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int64_t v[8] = {tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3};
        for (int c = 0; c < 8; c++) {
            // descale by CONST_BITS + PASS1_BITS + 3, level shift by 128, clamp -- through the 10-bit wrapped range table
            const int idx = (int)((v[c] + (1 << 17)) >> 18) & 1023;
            out[rrow * stride + c] = (unsigned char)(idx < 128 ? idx + 128 : idx < 512 ? 255 : idx < 896 ? 0 : idx - 896);
        }
    }
}

}  // namespace jpgdetail

struct JpegImage { int w = 0, h = 0; std::vector<unsigned char> rgb; };

inline JpegImage decode_jpeg_bytes(const std::vector<unsigned char> &b, const std::string &name) {
    using namespace jpgdetail;
    auto fail = [&name](const std::string &m) -> void { throw std::runtime_error("JPEG: " + m + ": " + name); };
    if (b.size() < 4 || b[0] != 0xFF || b[1] != 0xD8) fail("not a JPEG file");
    uint16_t qt[4][64]; bool haveQ[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    std::vector<Comp> comps;
    int W = 0, H = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restartInterval = 0;
    bool progressive = false, adobe = false; int adobeTransform = 0; bool jfif = false;
    size_t o = 2;
    bool done = false;
    while (!done && o + 4 <= b.size()) {
        if (b[o] != 0xFF) { o++; continue; }
        const int m = b[o + 1];
        if (m == 0xFF) { o++; continue; }
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) { o += 2; continue; }
        if (m == 0xD9) break;
        const size_t len = ((size_t)b[o + 2] << 8) | b[o + 3];
        if (len < 2 || o + 2 + len > b.size()) fail("truncated segment");
        const unsigned char *d = &b[o + 4];
        const size_t n = len - 2;
        if (m == 0xDB) {                                             // DQT
            for (size_t p = 0; p < n;) {
                const int pq = d[p] >> 4, tq = d[p] & 15; p++;
                if (tq > 3) fail("bad quantisation table id");
                for (int i = 0; i < 64; i++) {
                    const int v = pq ? ((d[p] << 8) | d[p + 1]) : d[p];
                    p += pq ? 2 : 1;
                    qt[tq][kZigzag[i]] = (uint16_t)v;
                }
                haveQ[tq] = true;
            }
        } else if (m == 0xC4) {                                      // DHT
            for (size_t p = 0; p < n;) {
                const int tc = d[p] >> 4, th = d[p] & 15; p++;
                if (th > 3 || tc > 1) fail("bad Huffman table id");
                Huff &h = tc ? ac[th] : dc[th];
                int total = 0;
                h.bits[0] = 0;
                for (int l = 1; l <= 16; l++) { h.bits[l] = d[p++]; total += h.bits[l]; }
                if (total > 256 || p + total > n) fail("bad Huffman table");
                memcpy(h.vals, d + p, total); p += total;
                h.present = true; h.build();
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {            // SOF0 / SOF1 / SOF2
            progressive = m == 0xC2;
            if (d[0] != 8) fail("only 8-bit samples are supported");
            H = (d[1] << 8) | d[2]; W = (d[3] << 8) | d[4];
            const int nc = d[5];
            if ((nc != 1 && nc != 3) || W <= 0 || H <= 0) fail("only 1- or 3-component images are supported");
            comps.resize(nc);
            for (int i = 0; i < nc; i++) {
                comps[i].id = d[6 + 3 * i]; comps[i].h = d[7 + 3 * i] >> 4; comps[i].v = d[7 + 3 * i] & 15; comps[i].tq = d[8 + 3 * i];
                if (comps[i].h < 1 || comps[i].v < 1 || comps[i].tq > 3) fail("bad component");
                if (comps[i].h > hmax) hmax = comps[i].h;
                if (comps[i].v > vmax) vmax = comps[i].v;
            }
            if (nc == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }
            mcux = (W + 8 * hmax - 1) / (8 * hmax); mcuy = (H + 8 * vmax - 1) / (8 * vmax);
            for (Comp &c : comps) {
                c.bw = mcux * c.h; c.bh = mcuy * c.v;
                c.dw = (W * c.h + hmax - 1) / hmax; c.dh = (H * c.v + vmax - 1) / vmax;
                c.cw = (c.dw + 7) / 8; c.chh = (c.dh + 7) / 8;
                c.coef.assign((size_t)c.bw * c.bh * 64, 0);
            }
        } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            fail("lossless / hierarchical / arithmetic-coded files are not supported");
        } else if (m == 0xDD) {
            restartInterval = (d[0] << 8) | d[1];
        } else if (m == 0xE0) {
            if (n >= 5 && !memcmp(d, "JFIF", 5)) jfif = true;
        } else if (m == 0xEE) {
            if (n >= 12 && !memcmp(d, "Adobe", 5)) { adobe = true; adobeTransform = d[11]; }
        } else if (m == 0xDA) {                                      // SOS + entropy-coded data
            if (comps.empty()) fail("scan before frame header");
            const int ns = d[0];
            if (ns < 1 || ns > (int)comps.size()) fail("bad scan");
            std::vector<Comp *> sc;
            for (int i = 0; i < ns; i++) {
                Comp *c = nullptr;
                for (Comp &k : comps) if (k.id == d[1 + 2 * i]) c = &k;
                if (!c) fail("scan names an unknown component");
                c->dcTab = d[2 + 2 * i] >> 4; c->acTab = d[2 + 2 * i] & 15;
                sc.push_back(c);
            }
            const int Ss = d[1 + 2 * ns], Se = d[2 + 2 * ns], Ah = d[3 + 2 * ns] >> 4, Al = d[3 + 2 * ns] & 15;
            if (!progressive && (Ss != 0 || Se != 63 || Ah != 0 || Al != 0)) fail("bad sequential scan parameters");
            if (progressive && (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1))) fail("bad progressive scan parameters");
            BitReader br(&b[o + 2 + len], b.data() + b.size());
            for (Comp *c : sc) c->pred = 0;
            int eobrun = 0;
            const bool interleaved = ns > 1;
            const int unitsX = interleaved ? mcux : sc[0]->cw, unitsY = interleaved ? mcuy : sc[0]->chh;
            int untilRestart = restartInterval;
            auto block_at = [](Comp *c, int bx, int by) { return &c->coef[((size_t)by * c->bw + bx) * 64]; };
            auto decode_block = [&](Comp *c, short *blk) {
                if (!progressive) {
                    const Huff &hd = dc[c->dcTab], &ha = ac[c->acTab];
                    if (!hd.present || !ha.present) fail("missing Huffman table");
                    int s = br.decode(hd);
                    if (s) { const int r = br.get(s); s = extend(r, s); }
                    c->pred += s; blk[0] = (short)c->pred;
                    for (int k = 1; k < 64; k++) {
                        int rs = br.decode(ha);
                        const int r = rs >> 4; s = rs & 15;
                        if (s) { k += r; if (k > 63) fail("corrupt coefficient run"); const int v = br.get(s); blk[kZigzag[k]] = (short)extend(v, s); }
                        else { if (r != 15) break; k += 15; }
                    }
                } else if (Ss == 0) {
                    if (Ah == 0) {
                        const Huff &hd = dc[c->dcTab];
                        if (!hd.present) fail("missing Huffman table");
                        int s = br.decode(hd);
                        if (s) { const int r = br.get(s); s = extend(r, s); }
                        c->pred += s; blk[0] = (short)(c->pred * (1 << Al));
                    } else if (br.get(1)) blk[0] |= (short)(1 << Al);
                } else if (Ah == 0) {
                    const Huff &ha = ac[c->acTab];
                    if (!ha.present) fail("missing Huffman table");
                    if (eobrun > 0) { eobrun--; return; }
                    for (int k = Ss; k <= Se; k++) {
                        const int rs = br.decode(ha);
                        const int r = rs >> 4; int s = rs & 15;
                        if (s) { k += r; if (k > 63) fail("corrupt coefficient run"); const int v = br.get(s); blk[kZigzag[k]] = (short)(extend(v, s) * (1 << Al)); }
                        else if (r == 15) k += 15;
                        else { eobrun = 1 << r; if (r) eobrun += br.get(r); eobrun--; break; }
                    }
                } else {
                    const Huff &ha = ac[c->acTab];
                    if (!ha.present) fail("missing Huffman table");
                    const int p1 = 1 << Al, m1 = -(1 << Al);
                    int k = Ss;
                    if (eobrun == 0) {
                        for (; k <= Se; k++) {
                            const int rs = br.decode(ha);
                            int r = rs >> 4, s = rs & 15;
                            if (s) s = br.get(1) ? p1 : m1;
                            else if (r != 15) { eobrun = 1 << r; if (r) eobrun += br.get(r); break; }
                            do {
                                short *co = blk + kZigzag[k];
                                if (*co != 0) {
                                    if (br.get(1) && (*co & p1) == 0) *co = (short)(*co + (*co >= 0 ? p1 : m1));
                                } else if (--r < 0) break;
                                k++;
                            } while (k <= Se);
                            if (s && k <= 63) blk[kZigzag[k]] = (short)s;
                        }
                    }
                    if (eobrun > 0) {
                        for (; k <= Se; k++) {
                            short *co = blk + kZigzag[k];
                            if (*co != 0 && br.get(1) && (*co & p1) == 0) *co = (short)(*co + (*co >= 0 ? p1 : m1));
                        }
                        eobrun--;
                    }
                }
            };
            for (int uy = 0; uy < unitsY; uy++)
                for (int ux = 0; ux < unitsX; ux++) {
                    if (restartInterval && untilRestart == 0) {
                        br.restart();
                        for (Comp *c : sc) c->pred = 0;
                        eobrun = 0; untilRestart = restartInterval;
                    }
                    if (interleaved) {
                        for (Comp *c : sc)
                            for (int by = 0; by < c->v; by++)
                                for (int bx = 0; bx < c->h; bx++) decode_block(c, block_at(c, ux * c->h + bx, uy * c->v + by));
                    } else decode_block(sc[0], block_at(sc[0], ux, uy));
                    untilRestart--;
                }
            // continue after the entropy-coded segment: the next marker that is not RSTn / a stuffed byte
            size_t q = (size_t)(br.p - b.data());
            while (q + 1 < b.size() && !(b[q] == 0xFF && b[q + 1] != 0x00 && !(b[q + 1] >= 0xD0 && b[q + 1] <= 0xD7) && b[q + 1] != 0xFF)) q++;
            o = q;
            continue;
        }
        o += 2 + len;
    }
    if (comps.empty()) fail("no frame header");
    // ---- dequantise + inverse DCT
    for (Comp &c : comps) {
        if (!haveQ[c.tq]) fail("missing quantisation table");
        const int stride = c.bw * 8;
        c.plane.assign((size_t)stride * c.bh * 8, 0);
        for (int by = 0; by < c.bh; by++)
            for (int bx = 0; bx < c.bw; bx++)
                idct_islow(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &c.plane[(size_t)by * 8 * stride + bx * 8], stride);
    }
    // ---- chroma upsampling to full resolution (triangle filter, edges replicated at the component's real extent)
    std::vector<std::vector<unsigned char>> full(comps.size());
    for (size_t ci = 0; ci < comps.size(); ci++) {
        Comp &c = comps[ci];
        const int stride = c.bw * 8;
        std::vector<unsigned char> &out = full[ci];
        out.assign((size_t)W * H, 0);
        if (c.h == hmax && c.v == vmax) {
            for (int y = 0; y < H; y++) memcpy(&out[(size_t)y * W], &c.plane[(size_t)y * stride], W);
        } else if (c.h * 2 == hmax && c.v == vmax) {                 // h2v1
            std::vector<unsigned char> row((size_t)c.dw * 2 + 2);
            for (int y = 0; y < H; y++) {
                const unsigned char *in = &c.plane[(size_t)y * stride];
                if (c.dw == 1) { row[0] = row[1] = in[0]; }
                else {
                    row[0] = in[0]; row[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
                    for (int x = 1; x < c.dw - 1; x++) {
                        row[2 * x] = (unsigned char)((in[x] * 3 + in[x - 1] + 1) >> 2);
                        row[2 * x + 1] = (unsigned char)((in[x] * 3 + in[x + 1] + 2) >> 2);
                    }
                    const int x = c.dw - 1;
                    row[2 * x] = (unsigned char)((in[x] * 3 + in[x - 1] + 1) >> 2); row[2 * x + 1] = in[x];
                }
                memcpy(&out[(size_t)y * W], row.data(), W);
            }
        } else if (c.h * 2 == hmax && c.v * 2 == vmax) {             // h2v2
            std::vector<unsigned char> row((size_t)c.dw * 2 + 2);
            for (int y = 0; y < H; y++) {
                const int iy = y >> 1;
                int ny = (y & 1) ? iy + 1 : iy - 1;                  // the farther of the two nearest input rows
                if (ny < 0) ny = 0;
                if (ny > c.dh - 1) ny = c.dh - 1;
                const unsigned char *in0 = &c.plane[(size_t)iy * stride], *in1 = &c.plane[(size_t)ny * stride];
                if (c.dw == 1) {
                    const int s = in0[0] * 3 + in1[0];
                    row[0] = (unsigned char)((s * 4 + 8) >> 4); row[1] = (unsigned char)((s * 4 + 7) >> 4);
                } else {
                    int thisc = in0[0] * 3 + in1[0], nextc = in0[1] * 3 + in1[1], lastc;
                    row[0] = (unsigned char)((thisc * 4 + 8) >> 4); row[1] = (unsigned char)((thisc * 3 + nextc + 7) >> 4);
                    lastc = thisc; thisc = nextc;
                    for (int x = 1; x < c.dw - 1; x++) {
                        nextc = in0[x + 1] * 3 + in1[x + 1];
                        row[2 * x] = (unsigned char)((thisc * 3 + lastc + 8) >> 4);
                        row[2 * x + 1] = (unsigned char)((thisc * 3 + nextc + 7) >> 4);
                        lastc = thisc; thisc = nextc;
                    }
                    const int x = c.dw - 1;
                    row[2 * x] = (unsigned char)((thisc * 3 + lastc + 8) >> 4); row[2 * x + 1] = (unsigned char)((thisc * 4 + 7) >> 4);
                }
                memcpy(&out[(size_t)y * W], row.data(), W);
            }
        } else fail("unsupported chroma subsampling (only 1x1, 2x1 and 2x2)");
    }
    // ---- colour conversion
    JpegImage im; im.w = W; im.h = H; im.rgb.resize((size_t)W * H * 3);
    if (comps.size() == 1) {
        for (size_t i = 0; i < (size_t)W * H; i++) im.rgb[3 * i] = im.rgb[3 * i + 1] = im.rgb[3 * i + 2] = full[0][i];
        return im;
    }
    bool ycc = true;
    if (adobe) ycc = adobeTransform != 0;
    else if (!jfif && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') ycc = false;
    if (!ycc) {
        for (size_t i = 0; i < (size_t)W * H; i++) { im.rgb[3 * i] = full[0][i]; im.rgb[3 * i + 1] = full[1][i]; im.rgb[3 * i + 2] = full[2][i]; }
        return im;
    }
    int crR[256], cbB[256]; int64_t crG[256], cbG[256];
    for (int i = 0; i < 256; i++) {                                 // 16-bit fixed point, JFIF coefficients
        const int64_t x = i - 128;
        crR[i] = (int)((91881 * x + 32768) >> 16);                   // 1.40200
        cbB[i] = (int)((116130 * x + 32768) >> 16);                  // 1.77200
        crG[i] = -46802 * x;                                         // 0.71414
        cbG[i] = -22554 * x + 32768;                                 // 0.34414
    }
    auto clamp8 = [](int v) { return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v); };
    for (size_t i = 0; i < (size_t)W * H; i++) {
        const int y = full[0][i], cb = full[1][i], cr = full[2][i];
        im.rgb[3 * i] = clamp8(y + crR[cr]);
        im.rgb[3 * i + 1] = clamp8(y + (int)((cbG[cb] + crG[cr]) >> 16));
        im.rgb[3 * i + 2] = clamp8(y + cbB[cb]);
    }
    return im;
}

}  // namespace lmc_host
