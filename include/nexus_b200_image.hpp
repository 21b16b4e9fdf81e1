// Image decoding for the C++ host layer (header-only, standard library only): PNG and JPEG into RGBA8, the two formats glTF
// embeds.  The reference decodes texture images with stb_image (src/Assets/IMGLoader.cpp:13-43: stbi_load(..., 4) -> RGBA8 pixels
// handed to AssetManager::AddTexture); an application that keeps stb feeds AddTexture directly and does not need this file.  The Python
// host layer uses Pillow for the same job (nexus_b200/gltf.py); tests/test_cpp_image.py compares the two decoders pixel for pixel.
//
//   nexus::DecodedImage img = nexus::DecodeImage(bytes, size);      // PNG: exact; JPEG: within the rounding of the IDCT / colour transform
//   uint32_t id = scene.GetAssetManager().AddTexture(img.rgba.data(), img.width, img.height, false, /*sRGB=*/true);
//
// PNG: every colour type and bit depth of the specification (16-bit samples keep their high byte), palette and colour-key transparency,
// Adam7 interlacing; its own inflate (RFC 1950 / 1951).  JPEG: baseline, extended sequential and PROGRESSIVE Huffman (SOF0 / SOF1 /
// SOF2: spectral selection and successive approximation, interleaved and per-component scans), 8-bit samples, greyscale or YCbCr
// with sampling factors up to 2 x 2, restart intervals, libjpeg-style triangle ("fancy") chroma upsampling; arithmetic-coded, lossless
// and CMYK files are rejected with an error that says so.  (Three of the reference's seven demo scenes embed progressive JPEGs.)
#ifndef NEXUS_B200_IMAGE_HPP
#define NEXUS_B200_IMAGE_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

namespace nexus {

struct DecodedImage { uint32_t width = 0, height = 0; std::vector<uint8_t> rgba; };   // rows top to bottom, 4 bytes per pixel
struct ImageError : std::runtime_error { using std::runtime_error::runtime_error; };

namespace imgdetail {

// ---------------------------------------------------------------------------------------------------- inflate ----
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint32_t buf = 0; int cnt = 0;
    uint32_t bits(int n)
    {
        while (cnt < n) { if (p >= end) throw ImageError("PNG: compressed data ends early"); buf |= (uint32_t)*p++ << cnt; cnt += 8; }
        const uint32_t v = buf & ((n == 32) ? 0xffffffffu : ((1u << n) - 1u));
        buf = n == 32 ? 0 : buf >> n; cnt -= n;
        return v;
    }
};
struct Huffman {                       // canonical code: symbols sorted by (length, value), count per length
    uint16_t count[16] = {0}; std::vector<uint16_t> symbol;
    void build(const uint8_t* lengths, int n)
    {
        std::memset(count, 0, sizeof(count)); symbol.assign((size_t)n, 0);
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; i++) if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader& br) const
    {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)br.bits(1);
            const int c = count[l];
            if (code - c < first) return symbol[(size_t)(index + (code - first))];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        throw ImageError("PNG: invalid Huffman code");
    }
};
inline std::vector<uint8_t> Inflate(const uint8_t* data, size_t n, size_t expected)
{
    if (n < 2 || (data[0] & 0x0f) != 8 || ((data[0] << 8) | data[1]) % 31) throw ImageError("PNG: not a zlib stream");
    if (data[1] & 0x20) throw ImageError("PNG: preset dictionaries are not allowed");
    BitReader br{data + 2, data + n};
    std::vector<uint8_t> out; out.reserve(expected);
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    bool last = false;
    while (!last) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {                                           // stored
            br.buf = 0; br.cnt = 0;
            if (br.end - br.p < 4) throw ImageError("PNG: compressed data ends early");
            const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
            br.p += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(br.end - br.p) < len) throw ImageError("PNG: corrupt stored block");
            out.insert(out.end(), br.p, br.p + len); br.p += len;
            continue;
        }
        if (type == 3) throw ImageError("PNG: invalid block type");
        Huffman lit, dist;
        uint8_t lengths[320];
        if (type == 1) {
            for (int i = 0; i < 144; i++) lengths[i] = 8;
            for (int i = 144; i < 256; i++) lengths[i] = 9;
            for (int i = 256; i < 280; i++) lengths[i] = 7;
            for (int i = 280; i < 288; i++) lengths[i] = 8;
            lit.build(lengths, 288);
            for (int i = 0; i < 30; i++) lengths[i] = 5;
            dist.build(lengths, 30);
        } else {
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncode; i++) cl[order[i]] = (uint8_t)br.bits(3);
            Huffman lencode; lencode.build(cl, 19);
            int idx = 0;
            while (idx < nlen + ndist) {
                const int sym = lencode.decode(br);
                if (sym < 16) { lengths[idx++] = (uint8_t)sym; continue; }
                int rep; uint8_t val = 0;
                if (sym == 16) { if (!idx) throw ImageError("PNG: corrupt code lengths"); val = lengths[idx - 1]; rep = 3 + (int)br.bits(2); }
                else if (sym == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (idx + rep > nlen + ndist) throw ImageError("PNG: corrupt code lengths");
                while (rep--) lengths[idx++] = val;
            }
            lit.build(lengths, nlen); dist.build(lengths + nlen, ndist);
        }
        while (true) {
            const int sym = lit.decode(br);
            if (sym < 256) { out.push_back((uint8_t)sym); continue; }
            if (sym == 256) break;
            if (sym > 285) throw ImageError("PNG: invalid length symbol");
            const int len = lbase[sym - 257] + (int)br.bits(lext[sym - 257]);
            const int ds = dist.decode(br);
            if (ds > 29) throw ImageError("PNG: invalid distance symbol");
            const size_t d = dbase[ds] + br.bits(dext[ds]);
            if (d > out.size()) throw ImageError("PNG: distance beyond the start of the data");
            const size_t from = out.size() - d;
            for (int i = 0; i < len; i++) out.push_back(out[from + (size_t)i]);
            if (out.size() > expected + 65536) throw ImageError("PNG: more image data than the header announces");
        }
    }
    return out;
}

// -------------------------------------------------------------------------------------------------------- PNG ----
inline uint32_t Be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int Paeth(int a, int b, int c) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); return pa <= pb && pa <= pc ? a : (pb <= pc ? b : c); }

inline DecodedImage DecodePNG(const uint8_t* data, size_t n)
{
    size_t off = 8;
    uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool haveHdr = false, done = false;
    while (!done && off + 12 <= n) {
        const uint32_t len = Be32(data + off); const uint8_t* type = data + off + 4; const uint8_t* body = data + off + 8;
        if (off + 12 + (size_t)len > n) throw ImageError("PNG: truncated chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw ImageError("PNG: short IHDR");
            w = Be32(body); h = Be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] || body[11] || interlace > 1) throw ImageError("PNG: unknown compression, filter or interlace method");
            haveHdr = true;
        }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) done = true;
        off += 12 + (size_t)len;
    }
    if (!haveHdr || !w || !h || idat.empty()) throw ImageError("PNG: missing header or image data");
    if ((uint64_t)w * h > (1ull << 28)) throw ImageError("PNG: image too large");
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depthOk = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                       : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (!channels || !depthOk) throw ImageError("PNG: invalid colour type / bit depth");
    if (ctype == 3 && palette.size() < 3) throw ImageError("PNG: palette image without a palette");
    const int bitsPerPixel = channels * depth, bpp = std::max(1, bitsPerPixel / 8);
    auto rowBytes = [&](uint32_t pw) { return ((size_t)pw * (size_t)bitsPerPixel + 7) / 8; };
    // passes: the whole image, or the seven Adam7 sub-images
    struct Pass { uint32_t x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    const Pass whole[1] = {{0, 0, 1, 1}};
    const Pass* passes = interlace ? adam7 : whole; const int nPasses = interlace ? 7 : 1;
    size_t expected = 0;
    for (int k = 0; k < nPasses; k++) {
        const uint32_t pw = (w - passes[k].x0 + passes[k].dx - 1) / passes[k].dx, ph = (h - passes[k].y0 + passes[k].dy - 1) / passes[k].dy;
        if (w > passes[k].x0 && h > passes[k].y0 && pw && ph) expected += (rowBytes(pw) + 1) * ph;
    }
    const std::vector<uint8_t> raw = Inflate(idat.data(), idat.size(), expected);
    if (raw.size() < expected) throw ImageError("PNG: image data ends early");

    DecodedImage img; img.width = w; img.height = h; img.rgba.assign((size_t)4 * w * h, 255);
    auto sample = [&](const uint8_t* row, uint32_t x, int c) -> uint32_t {      // sample c of pixel x as stored (not scaled)
        if (depth == 8) return row[(size_t)x * channels + c];
        if (depth == 16) return row[((size_t)x * channels + c) * 2];          // high byte
        const uint32_t bit = x * (uint32_t)depth; const uint32_t v = row[bit >> 3] >> (8 - depth - (bit & 7)); return v & ((1u << depth) - 1u);
    };
    // colour-key transparency (types 0 and 2): compared on the full stored sample values
    auto sample16 = [&](const uint8_t* row, uint32_t x, int c) -> uint32_t { return depth == 16 ? ((uint32_t)row[((size_t)x * channels + c) * 2] << 8) | row[((size_t)x * channels + c) * 2 + 1] : sample(row, x, c); };
    size_t pos = 0;
    std::vector<uint8_t> prev, cur;
    for (int k = 0; k < nPasses; k++) {
        if (w <= passes[k].x0 || h <= passes[k].y0) continue;
        const uint32_t pw = (w - passes[k].x0 + passes[k].dx - 1) / passes[k].dx, ph = (h - passes[k].y0 + passes[k].dy - 1) / passes[k].dy;
        if (!pw || !ph) continue;
        const size_t rb = rowBytes(pw);
        prev.assign(rb, 0); cur.assign(rb, 0);
        for (uint32_t y = 0; y < ph; y++) {
            const uint8_t filter = raw[pos++]; const uint8_t* src = raw.data() + pos; pos += rb;
            for (size_t i = 0; i < rb; i++) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int v = src[i];
                switch (filter) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += Paeth(a, b, c); break;
                                  default: throw ImageError("PNG: invalid filter type"); }
                cur[i] = (uint8_t)v;
            }
            const uint32_t oy = passes[k].y0 + y * passes[k].dy;
            for (uint32_t x = 0; x < pw; x++) {
                uint8_t* o = img.rgba.data() + 4 * ((size_t)oy * w + passes[k].x0 + (size_t)x * passes[k].dx);
                if (ctype == 3) {
                    const uint32_t i = sample(cur.data(), x, 0);
                    if (3 * (size_t)i + 2 >= palette.size()) throw ImageError("PNG: palette index out of range");
                    o[0] = palette[3 * i]; o[1] = palette[3 * i + 1]; o[2] = palette[3 * i + 2]; o[3] = i < trns.size() ? trns[i] : 255;
                } else if (ctype == 0 || ctype == 4) {
                    uint32_t g = sample(cur.data(), x, 0);
                    if (depth < 8) g = g * 255u / ((1u << depth) - 1u);
                    o[0] = o[1] = o[2] = (uint8_t)g;
                    o[3] = ctype == 4 ? (uint8_t)sample(cur.data(), x, 1) : 255;
                    if (ctype == 0 && trns.size() >= 2 && sample16(cur.data(), x, 0) == (((uint32_t)trns[0] << 8) | trns[1])) o[3] = 0;
                } else {
                    o[0] = (uint8_t)sample(cur.data(), x, 0); o[1] = (uint8_t)sample(cur.data(), x, 1); o[2] = (uint8_t)sample(cur.data(), x, 2);
                    o[3] = ctype == 6 ? (uint8_t)sample(cur.data(), x, 3) : 255;
                    if (ctype == 2 && trns.size() >= 6 && sample16(cur.data(), x, 0) == (((uint32_t)trns[0] << 8) | trns[1]) &&
                        sample16(cur.data(), x, 1) == (((uint32_t)trns[2] << 8) | trns[3]) && sample16(cur.data(), x, 2) == (((uint32_t)trns[4] << 8) | trns[5])) o[3] = 0;
                }
            }
            prev.swap(cur);
        }
    }
    return img;
}

// ------------------------------------------------------------------------------------------------------- JPEG ----
struct JpegHuff { uint8_t bits[17] = {0}; uint8_t vals[256] = {0}; int mincode[17], maxcode[18], valptr[17]; bool present = false;
    void build()
    {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) { valptr[l] = k; mincode[l] = code; code += bits[l]; k += bits[l]; maxcode[l] = bits[l] ? code - 1 : -1; code <<= 1; }
        maxcode[17] = 0x7fffffff; present = true;
    }
};
struct JpegBits {
    const uint8_t* p; const uint8_t* end; uint32_t buf = 0; int cnt = 0; bool hitMarker = false;
    void fill()
    {
        while (cnt <= 24) {
            uint32_t b = 0;
            if (!hitMarker && p < end) {
                b = *p;
                if (b == 0xff) { if (p + 1 < end && p[1] == 0) p += 2; else { hitMarker = true; b = 0; } }
                else p++;
            }
            buf |= b << (24 - cnt); cnt += 8;
        }
    }
    int get(int n) { if (!n) return 0; if (cnt < n) fill(); const int v = (int)(buf >> (32 - n)); buf <<= n; cnt -= n; return v; }
    int decode(const JpegHuff& h)
    {
        int code = 0;
        for (int l = 1; l <= 16; l++) { code = (code << 1) | get(1); if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]]; }
        throw ImageError("JPEG: invalid Huffman code");
    }
    static int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }
    void reset() { buf = 0; cnt = 0; hitMarker = false; }
};
inline void Idct8x8(const float* in, uint8_t* out, size_t stride)
{
    // separable floating-point inverse DCT (the definition, with the cosines tabulated), + 128, rounded and clamped
    static float c[8][8]; static bool init = false;
    if (!init) { for (int x = 0; x < 8; x++) for (int u = 0; u < 8; u++) c[x][u] = (u ? 1.0f : 0.70710678118654752f) * 0.5f * (float)std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0); init = true; }
    float tmp[64];
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) { float s = 0; for (int u = 0; u < 8; u++) s += c[x][u] * in[8 * y + u]; tmp[8 * y + x] = s; }
    for (int x = 0; x < 8; x++) for (int y = 0; y < 8; y++) {
        float s = 0; for (int v = 0; v < 8; v++) s += c[y][v] * tmp[8 * v + x];
        const int r = (int)std::lround(s + 128.0f);
        out[(size_t)y * stride + x] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}
// Baseline, extended-sequential and progressive Huffman JPEG (ITU T.81): every scan is decoded into per-component coefficient arrays
// (sequential files fill them in one go, progressive files by spectral selection and successive approximation over many scans), and
// the image is reconstructed once at the end: dequantisation, inverse DCT, chroma upsampling, YCbCr -> RGB.
inline DecodedImage DecodeJPEG(const uint8_t* data, size_t n)
{
    static const uint8_t zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; uint32_t bw = 0, bh = 0, cw = 0, ch = 0; std::vector<int16_t> coef; std::vector<uint8_t> plane; };
    float qt[4][64]; bool haveQ[4] = {false, false, false, false};
    JpegHuff dc[4], ac[4];
    Comp comp[3]; int nComp = 0; uint32_t w = 0, h = 0; int restart = 0; bool haveFrame = false, progressive = false, sawScan = false;
    int hmax = 1, vmax = 1; uint32_t mx = 0, my = 0;
    size_t off = 2;
    auto be16 = [&](size_t o) { if (o + 2 > n) throw ImageError("JPEG: truncated file"); return (int)((data[o] << 8) | data[o + 1]); };
    while (off + 4 <= n) {
        if (data[off] != 0xff) { off++; continue; }
        const int marker = data[off + 1];
        if (marker == 0xff) { off++; continue; }
        off += 2;
        if (marker == 0xd8 || marker == 0x01 || (marker >= 0xd0 && marker <= 0xd7)) continue;
        if (marker == 0xd9) break;
        const int len = be16(off);
        if (len < 2 || off + (size_t)len > n) throw ImageError("JPEG: truncated segment");
        const uint8_t* s = data + off + 2; const int sl = len - 2;
        if (marker == 0xdb) {                                          // DQT
            for (int i = 0; i < sl;) {
                const int pq = s[i] >> 4, tq = s[i] & 15; i++;
                if (tq > 3 || i + (pq ? 128 : 64) > sl) throw ImageError("JPEG: corrupt quantisation table");
                for (int k = 0; k < 64; k++) { qt[tq][zigzag[k]] = pq ? (float)((s[i] << 8) | s[i + 1]) : (float)s[i]; i += pq ? 2 : 1; }
                haveQ[tq] = true;
            }
        } else if (marker == 0xc4) {                                   // DHT
            for (int i = 0; i < sl;) {
                const int tc = s[i] >> 4, th = s[i] & 15; i++;
                if (tc > 1 || th > 3 || i + 16 > sl) throw ImageError("JPEG: corrupt Huffman table");
                JpegHuff& t = tc ? ac[th] : dc[th];
                int total = 0; for (int l = 1; l <= 16; l++) { t.bits[l] = s[i + l - 1]; total += t.bits[l]; }
                i += 16;
                if (total > 256 || i + total > sl) throw ImageError("JPEG: corrupt Huffman table");
                std::memcpy(t.vals, s + i, (size_t)total); i += total;
                t.build();
            }
        } else if (marker == 0xc0 || marker == 0xc1 || marker == 0xc2) {   // SOF0 / SOF1 / SOF2
            if (haveFrame) throw ImageError("JPEG: more than one frame");
            if (sl < 6 || s[0] != 8) throw ImageError("JPEG: only 8-bit samples are supported");
            progressive = marker == 0xc2;
            h = (uint32_t)((s[1] << 8) | s[2]); w = (uint32_t)((s[3] << 8) | s[4]); nComp = s[5];
            if ((nComp != 1 && nComp != 3) || sl < 6 + 3 * nComp || !w || !h) throw ImageError("JPEG: unsupported component count or empty image");
            if ((uint64_t)w * h > (1ull << 28)) throw ImageError("JPEG: image too large");
            for (int k = 0; k < nComp; k++) { comp[k].id = s[6 + 3 * k]; comp[k].h = s[7 + 3 * k] >> 4; comp[k].v = s[7 + 3 * k] & 15; comp[k].tq = s[8 + 3 * k];
                if (comp[k].h < 1 || comp[k].h > 2 || comp[k].v < 1 || comp[k].v > 2 || comp[k].tq > 3) throw ImageError("JPEG: unsupported sampling factors"); }
            if (nComp == 1) comp[0].h = comp[0].v = 1;                  // a single component is never interleaved: 8 x 8 MCUs
            for (int k = 0; k < nComp; k++) { hmax = std::max(hmax, comp[k].h); vmax = std::max(vmax, comp[k].v); }
            mx = (w + 8u * hmax - 1) / (8u * hmax); my = (h + 8u * vmax - 1) / (8u * vmax);
            for (int k = 0; k < nComp; k++) {
                Comp& c = comp[k];
                c.bw = mx * (uint32_t)c.h; c.bh = my * (uint32_t)c.v;
                c.cw = (w * (uint32_t)c.h + hmax - 1) / hmax; c.ch = (h * (uint32_t)c.v + vmax - 1) / vmax;
                c.coef.assign((size_t)64 * c.bw * c.bh, 0);
            }
            haveFrame = true;
        } else if (marker >= 0xc3 && marker <= 0xcf && marker != 0xc4 && marker != 0xc8 && marker != 0xcc) throw ImageError("JPEG: lossless / hierarchical / arithmetic-coded files are not supported");
        else if (marker == 0xdd) { if (sl < 2) throw ImageError("JPEG: corrupt DRI"); restart = (s[0] << 8) | s[1]; }
        else if (marker == 0xda) {                                     // SOS: one scan
            if (!haveFrame) throw ImageError("JPEG: scan before the frame header");
            const int ns = sl >= 1 ? s[0] : 0;
            if (ns < 1 || ns > nComp || sl < 4 + 2 * ns) throw ImageError("JPEG: corrupt scan header");
            int order[3];
            for (int k = 0; k < ns; k++) {
                int ci = -1; for (int j = 0; j < nComp; j++) if (comp[j].id == s[1 + 2 * k]) ci = j;
                if (ci < 0) throw ImageError("JPEG: scan names an unknown component");
                comp[ci].td = s[2 + 2 * k] >> 4; comp[ci].ta = s[2 + 2 * k] & 15;
                if (comp[ci].td > 3 || comp[ci].ta > 3) throw ImageError("JPEG: corrupt scan header");
                order[k] = ci;
            }
            const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
            if (progressive) { if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13 || Ah > 13) throw ImageError("JPEG: invalid progressive scan parameters"); }
            else if (Ss != 0 || Se != 63 || Ah != 0 || Al != 0) throw ImageError("JPEG: invalid sequential scan parameters");
            for (int k = 0; k < ns; k++) {
                const Comp& c = comp[order[k]];
                const bool needDc = !progressive || (Ss == 0 && Ah == 0), needAc = !progressive || Ss > 0;
                if ((needDc && !dc[c.td].present) || (needAc && !ac[c.ta].present) || !haveQ[c.tq]) throw ImageError("JPEG: scan uses a table the file does not define");
            }
            // scan geometry: one component -> its own block grid in raster order; several -> MCUs of h x v blocks per component
            const bool inter = ns > 1;
            const Comp& c0 = comp[order[0]];
            const uint32_t unitsX = inter ? mx : (c0.cw + 7) / 8, unitsY = inter ? my : (c0.ch + 7) / 8;
            JpegBits br{data + off + (size_t)len, data + n};
            int untilRestart = restart, eobrun = 0;
            for (int k = 0; k < ns; k++) comp[order[k]].pred = 0;
            auto refine = [&](int16_t& cf, int p1, int m1) { if (br.get(1) && (cf & p1) == 0) cf = (int16_t)(cf + (cf >= 0 ? p1 : m1)); };
            auto block = [&](Comp& c, int16_t* b) {
                if (!progressive) {
                    const int t = br.decode(dc[c.td]);
                    if (t > 11) throw ImageError("JPEG: corrupt DC coefficient");
                    c.pred += t ? JpegBits::extend(br.get(t), t) : 0;
                    b[0] = (int16_t)c.pred;
                    for (int i = 1; i < 64;) {
                        const int rs = br.decode(ac[c.ta]), r = rs >> 4, sz = rs & 15;
                        if (!sz) { if (r == 15) { i += 16; continue; } break; }
                        i += r;
                        if (i > 63) throw ImageError("JPEG: corrupt AC coefficients");
                        b[zigzag[i]] = (int16_t)JpegBits::extend(br.get(sz), sz);
                        i++;
                    }
                } else if (Ss == 0) {
                    if (Ah == 0) {                                      // DC, first pass
                        const int t = br.decode(dc[c.td]);
                        if (t > 11) throw ImageError("JPEG: corrupt DC coefficient");
                        c.pred += t ? JpegBits::extend(br.get(t), t) : 0;
                        b[0] = (int16_t)(c.pred * (1 << Al));
                    } else if (br.get(1)) b[0] = (int16_t)(b[0] | (1 << Al));   // DC, refinement: one more bit
                } else if (Ah == 0) {                                   // AC band, first pass
                    if (eobrun > 0) { eobrun--; return; }
                    for (int k = Ss; k <= Se;) {
                        const int rs = br.decode(ac[c.ta]), r = rs >> 4, sz = rs & 15;
                        if (!sz) {
                            if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += br.get(r); break; }
                            k += 16;
                        } else {
                            k += r;
                            if (k > Se) throw ImageError("JPEG: corrupt AC coefficients");
                            b[zigzag[k]] = (int16_t)(JpegBits::extend(br.get(sz), sz) * (1 << Al));
                            k++;
                        }
                    }
                } else {                                                // AC band, refinement (T.81 G.1.2.3)
                    const int p1 = 1 << Al, m1 = -(1 << Al);
                    int k = Ss;
                    if (eobrun == 0) {
                        for (; k <= Se; k++) {
                            const int rs = br.decode(ac[c.ta]); int r = rs >> 4; const int sz = rs & 15;
                            int val = 0;
                            if (sz) { if (sz != 1) throw ImageError("JPEG: corrupt AC refinement"); val = br.get(1) ? p1 : m1; }
                            else if (r < 15) { eobrun = 1 << r; if (r) eobrun += br.get(r); break; }
                            // advance over r zero-history coefficients, refining the nonzero ones on the way
                            for (; k <= Se; k++) {
                                int16_t& cf = b[zigzag[k]];
                                if (cf != 0) refine(cf, p1, m1);
                                else if (--r < 0) break;
                            }
                            if (sz && k <= Se) b[zigzag[k]] = (int16_t)val;
                        }
                    }
                    if (eobrun > 0) {
                        for (; k <= Se; k++) { int16_t& cf = b[zigzag[k]]; if (cf != 0) refine(cf, p1, m1); }
                        eobrun--;
                    }
                }
            };
            for (uint32_t uy = 0; uy < unitsY; uy++) for (uint32_t ux = 0; ux < unitsX; ux++) {
                if (restart && untilRestart == 0) {
                    // the restart marker: byte aligned, right where the bit reader stopped
                    br.reset();
                    while (br.p + 1 < br.end && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) br.p++;
                    if (br.p + 1 < br.end) br.p += 2;
                    for (int k = 0; k < ns; k++) comp[order[k]].pred = 0;
                    eobrun = 0;
                    untilRestart = restart;
                }
                if (inter) {
                    for (int k = 0; k < ns; k++) { Comp& c = comp[order[k]];
                        for (int v = 0; v < c.v; v++) for (int hh = 0; hh < c.h; hh++)
                            block(c, c.coef.data() + 64 * ((size_t)(uy * c.v + v) * c.bw + (size_t)(ux * c.h + hh))); }
                } else { Comp& c = comp[order[0]]; block(c, c.coef.data() + 64 * ((size_t)uy * c.bw + ux)); }
                if (restart) untilRestart--;
            }
            sawScan = true;
            // the entropy-coded data follows the header: skip to the next marker that is not a restart marker or a stuffed byte
            size_t q = off + (size_t)len;
            while (q + 1 < n && !(data[q] == 0xff && data[q + 1] != 0 && !(data[q + 1] >= 0xd0 && data[q + 1] <= 0xd7) && data[q + 1] != 0xff)) q++;
            off = q;
            continue;
        }
        off += (size_t)len;
    }
    if (!haveFrame || !sawScan) throw ImageError("JPEG: the file has no image scan");

    // ---- reconstruction: dequantise + inverse DCT per block
    float blockf[64];
    for (int k = 0; k < nComp; k++) {
        Comp& c = comp[k];
        const uint32_t pw = c.bw * 8u, ph = c.bh * 8u;
        c.plane.assign((size_t)pw * ph, 0);
        for (uint32_t by = 0; by < c.bh; by++) for (uint32_t bx = 0; bx < c.bw; bx++) {
            const int16_t* b = c.coef.data() + 64 * ((size_t)by * c.bw + bx);
            for (int i = 0; i < 64; i++) blockf[i] = (float)b[i] * qt[c.tq][i];
            Idct8x8(blockf, c.plane.data() + ((size_t)by * 8) * pw + (size_t)bx * 8, pw);
        }
    }
    // ---- to RGBA: chroma upsampled with the triangle filter libjpeg calls "fancy upsampling" (3/4, 1/4 per axis), replication otherwise
    DecodedImage img; img.width = w; img.height = h; img.rgba.assign((size_t)4 * w * h, 255);
    auto at = [&](const Comp& c, int fx, int fy, uint32_t x, uint32_t y) -> float {      // component value at full-resolution pixel (x, y)
        const uint32_t pw = c.bw * 8u;
        auto px = [&](long cx, long cy) { cx = cx < 0 ? 0 : cx >= (long)c.cw ? (long)c.cw - 1 : cx; cy = cy < 0 ? 0 : cy >= (long)c.ch ? (long)c.ch - 1 : cy;
                                           return (float)c.plane[(size_t)cy * pw + (size_t)cx]; };
        if (fx == 1 && fy == 1) return px((long)x, (long)y);
        const long cx = (long)(x / (uint32_t)fx), cy = (long)(y / (uint32_t)fy);
        const long nx = fx == 2 ? ((x & 1u) ? cx + 1 : cx - 1) : cx, ny = fy == 2 ? ((y & 1u) ? cy + 1 : cy - 1) : cy;
        const float wx = fx == 2 ? 0.25f : 0.0f, wy = fy == 2 ? 0.25f : 0.0f;
        return (1 - wy) * ((1 - wx) * px(cx, cy) + wx * px(nx, cy)) + wy * ((1 - wx) * px(cx, ny) + wx * px(nx, ny));
    };
    auto clamp8 = [](float v) { const int r = (int)std::lround(v); return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r); };
    for (uint32_t y = 0; y < h; y++) for (uint32_t x = 0; x < w; x++) {
        uint8_t* o = img.rgba.data() + 4 * ((size_t)y * w + x);
        if (nComp == 1) { o[0] = o[1] = o[2] = comp[0].plane[(size_t)y * comp[0].bw * 8u + x]; continue; }
        const float Y = at(comp[0], hmax / comp[0].h, vmax / comp[0].v, x, y), Cb = at(comp[1], hmax / comp[1].h, vmax / comp[1].v, x, y) - 128.0f,
                    Cr = at(comp[2], hmax / comp[2].h, vmax / comp[2].v, x, y) - 128.0f;
        o[0] = clamp8(Y + 1.402f * Cr); o[1] = clamp8(Y - 0.344136f * Cb - 0.714136f * Cr); o[2] = clamp8(Y + 1.772f * Cb);
    }
    return img;
}

}  // namespace imgdetail

// PNG or JPEG bytes -> RGBA8 (what stbi_load_from_memory(..., 4) gives the reference, IMGLoader.cpp:13-43)
inline DecodedImage DecodeImage(const unsigned char* data, size_t size)
{
    static const unsigned char pngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (size >= 8 && !std::memcmp(data, pngSig, 8)) return imgdetail::DecodePNG(data, size);
    if (size >= 4 && data[0] == 0xff && data[1] == 0xd8) return imgdetail::DecodeJPEG(data, size);
    throw ImageError("image: neither a PNG nor a JPEG file");
}
inline DecodedImage LoadImageFile(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ImageError("cannot open " + path);
    const std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return DecodeImage(bytes.data(), bytes.size());
}

}  // namespace nexus
#endif /* NEXUS_B200_IMAGE_HPP */
