// ImageIO.cpp -- see ImageIO.h.
#include "ImageIO.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace vr {

namespace {

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}

void put32be(std::vector<uint8_t>& v, uint32_t x)
{
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}

void chunk(std::vector<uint8_t>& png, const char type[4], const std::vector<uint8_t>& data)
{
    put32be(png, (uint32_t)data.size());
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    const uint32_t crc = crc32_update(0xffffffffu, png.data() + start, png.size() - start) ^ 0xffffffffu;
    put32be(png, crc);
}

bool dump(const std::string& fn, const std::vector<uint8_t>& bytes)
{
    FILE* f = std::fopen(fn.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    std::fclose(f);
    return ok;
}

}  // namespace

// ------------------------------------------------------------------ baseline JPEG (JFIF)
// RendererCore.cpp:176 asks stb_image_write for quality 100, i.e. quantisation tables of all ones and no
// chroma subsampling.  This writer does the same (8x8 float DCT, round to nearest, zig-zag, DC
// differences, AC run lengths) and, instead of the default Annex K tables, emits Huffman tables
// optimised for the image (ITU-T T.81 K.2: two passes over the symbols), which every baseline decoder
// accepts.  Only the pixels matter to the caller, not the byte stream.
namespace {

const uint8_t kZigZag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
                             13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59,
                             52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    uint8_t bits[17] = {0};          // bits[l] = number of codes of length l
    std::vector<uint8_t> vals;       // symbols in order of increasing code length
    uint16_t code[256] = {0};
    uint8_t size[256] = {0};
};

// T.81 annex K.2: code lengths from symbol frequencies, limited to 16 bits, no all-ones code
void buildHuffman(const uint32_t freq_in[256], HuffTable& t)
{
    long freq[257];
    int codesize[257] = {0}, others[257];
    for (int i = 0; i < 256; ++i) freq[i] = freq_in[i];
    freq[256] = 1;                                   // reserves the all-ones code
    for (int i = 0; i < 257; ++i) others[i] = -1;
    for (;;) {
        int c1 = -1, c2 = -1;
        long v = 0;
        for (int i = 0; i < 257; ++i) if (freq[i] && (c1 < 0 || freq[i] <= v)) { v = freq[i]; c1 = i; }
        v = 0;
        for (int i = 0; i < 257; ++i) if (freq[i] && i != c1 && (c2 < 0 || freq[i] <= v)) { v = freq[i]; c2 = i; }
        if (c2 < 0) break;
        freq[c1] += freq[c2];
        freq[c2] = 0;
        for (++codesize[c1]; others[c1] >= 0; ++codesize[c1]) c1 = others[c1];
        others[c1] = c2;
        for (++codesize[c2]; others[c2] >= 0; ++codesize[c2]) c2 = others[c2];
    }
    int bits[64] = {0};
    for (int i = 0; i < 257; ++i) if (codesize[i]) ++bits[codesize[i]];
    for (int i = 63; i > 16; --i) {                  // K.3: fold lengths above 16 back
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) --j;
            bits[i] -= 2; ++bits[i - 1];
            bits[j + 1] += 2; --bits[j];
        }
    }
    int i = 16;
    while (bits[i] == 0) --i;
    --bits[i];                                       // drop the reserved symbol
    for (int l = 1; l <= 16; ++l) t.bits[l] = (uint8_t)bits[l];
    t.vals.clear();
    for (int l = 1; l <= 63; ++l)
        for (int s = 0; s < 256; ++s) if (codesize[s] == l) t.vals.push_back((uint8_t)s);
    // canonical codes in the order of `vals` (lengths re-read from `bits`, which the folding changed)
    uint16_t code = 0;
    size_t k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int n = 0; n < t.bits[l]; ++n, ++k) { t.code[t.vals[k]] = code++; t.size[t.vals[k]] = (uint8_t)l; }
        code <<= 1;
    }
}

struct JpegBits {
    std::vector<uint8_t>& out;
    uint32_t acc = 0;
    int n = 0;
    void put(uint32_t code, int len)
    {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            out.push_back(b);
            if (b == 0xff) out.push_back(0);         // byte stuffing
            n -= 8;
        }
    }
    void flush() { if (n) put(0x7f, 8 - n); }        // pad with ones
};

inline int magnitudeBits(int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; }

void putMarker(std::vector<uint8_t>& o, uint8_t m, const std::vector<uint8_t>& payload)
{
    o.push_back(0xff); o.push_back(m);
    const size_t len = payload.size() + 2;
    o.push_back((uint8_t)(len >> 8)); o.push_back((uint8_t)len);
    o.insert(o.end(), payload.begin(), payload.end());
}

}  // namespace

bool writeJPG(const std::string& fn, int w, int h, const uint8_t* rgb)
{
    if (w < 1 || h < 1 || w > 65535 || h > 65535 || !rgb) return false;
    const int bw = (w + 7) / 8, bh = (h + 7) / 8;
    // DCT basis: c[u][x] = alpha(u)/2 * cos((2x+1) u pi / 16)
    float basis[8][8];
    for (int u = 0; u < 8; ++u)
        for (int x = 0; x < 8; ++x)
            basis[u][x] = (float)((u == 0 ? std::sqrt(0.5) : 1.0) * 0.5 * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    // quantised coefficients of every block, component-major inside a block triple, zig-zag order
    std::vector<int16_t> coef((size_t)bw * bh * 3 * 64);
    for (int by = 0; by < bh; ++by)
        for (int bx = 0; bx < bw; ++bx) {
            float px[3][64];
            for (int y = 0; y < 8; ++y)
                for (int x = 0; x < 8; ++x) {
                    const int sx = std::min(bx * 8 + x, w - 1), sy = std::min(by * 8 + y, h - 1);   // edge replication
                    const uint8_t* p = rgb + ((size_t)sy * w + sx) * 3;
                    const float r = p[0], g = p[1], b = p[2];
                    px[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
                    px[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                    px[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                }
            for (int c = 0; c < 3; ++c) {
                float tmp[64], dct[64];
                for (int y = 0; y < 8; ++y)
                    for (int u = 0; u < 8; ++u) {
                        float s = 0;
                        for (int x = 0; x < 8; ++x) s += px[c][y * 8 + x] * basis[u][x];
                        tmp[y * 8 + u] = s;
                    }
                for (int v = 0; v < 8; ++v)
                    for (int u = 0; u < 8; ++u) {
                        float s = 0;
                        for (int y = 0; y < 8; ++y) s += tmp[y * 8 + u] * basis[v][y];
                        dct[v * 8 + u] = s;
                    }
                int16_t* dst = &coef[(((size_t)by * bw + bx) * 3 + c) * 64];
                for (int k = 0; k < 64; ++k) {
                    const float q = dct[kZigZag[k]];                    // quantiser step 1 (quality 100)
                    int v = (int)std::lround(q);
                    dst[k] = (int16_t)std::max(-1023, std::min(1023, v));
                }
            }
        }
    // symbols of one block: DC category, then (run << 4 | size) pairs; visit() sees (table class, symbol, extra bits, nbits)
    auto walk = [&](auto&& visit) {
        int pred[3] = {0, 0, 0};
        for (size_t b = 0; b < (size_t)bw * bh; ++b)
            for (int c = 0; c < 3; ++c) {
                const int16_t* q = &coef[(b * 3 + c) * 64];
                const int chroma = c ? 1 : 0;
                const int diff = q[0] - pred[c];
                pred[c] = q[0];
                int nb = magnitudeBits(diff);
                visit(chroma * 2 + 0, nb, diff < 0 ? diff - 1 : diff, nb);
                int last = 63;
                while (last > 0 && q[last] == 0) --last;
                int run = 0;
                for (int k = 1; k <= last; ++k) {
                    if (q[k] == 0) { ++run; continue; }
                    while (run > 15) { visit(chroma * 2 + 1, 0xf0, 0, 0); run -= 16; }
                    nb = magnitudeBits(q[k]);
                    visit(chroma * 2 + 1, (run << 4) | nb, q[k] < 0 ? q[k] - 1 : q[k], nb);
                    run = 0;
                }
                if (last < 63) visit(chroma * 2 + 1, 0x00, 0, 0);     // end of block
            }
    };
    uint32_t freq[4][256];
    std::memset(freq, 0, sizeof freq);
    walk([&](int table, int symbol, int, int) { ++freq[table][symbol]; });
    HuffTable ht[4];                                                    // 0 DC luma, 1 AC luma, 2 DC chroma, 3 AC chroma
    for (int t = 0; t < 4; ++t) buildHuffman(freq[t], ht[t]);

    std::vector<uint8_t> o;
    o.push_back(0xff); o.push_back(0xd8);                               // SOI
    putMarker(o, 0xe0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; ++t) {                                       // DQT: all ones
        std::vector<uint8_t> q(65, 1);
        q[0] = (uint8_t)t;
        putMarker(o, 0xdb, q);
    }
    {
        std::vector<uint8_t> sof = {8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3,
                                    1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1};
        putMarker(o, 0xc0, sof);                                        // SOF0, 4:4:4
    }
    for (int t = 0; t < 4; ++t) {                                       // DHT
        std::vector<uint8_t> d;
        d.push_back((uint8_t)(((t & 1) << 4) | (t >> 1)));              // class (0 DC / 1 AC) << 4 | destination
        for (int l = 1; l <= 16; ++l) d.push_back(ht[t].bits[l]);
        d.insert(d.end(), ht[t].vals.begin(), ht[t].vals.end());
        putMarker(o, 0xc4, d);
    }
    putMarker(o, 0xda, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});       // SOS
    JpegBits bits{o};
    walk([&](int table, int symbol, int extra, int nbits) {
        bits.put(ht[table].code[symbol], ht[table].size[symbol]);
        if (nbits) bits.put((uint32_t)extra, nbits);
    });
    bits.flush();
    o.push_back(0xff); o.push_back(0xd9);                               // EOI
    return dump(fn, o);
}

bool writePNG(const std::string& fn, int w, int h, const uint8_t* rgb)
{
    if (w < 1 || h < 1 || !rgb) return false;
    // raw scanlines: filter byte 0 + RGB
    std::vector<uint8_t> raw;
    raw.reserve((size_t)h * (1 + (size_t)w * 3));
    for (int y = 0; y < h; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
    }
    // zlib container with stored (uncompressed) deflate blocks
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    while (pos < raw.size() || raw.empty()) {
        const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
        const bool last = pos + n == raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(n & 0xff); z.push_back(n >> 8);
        z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        for (size_t i = 0; i < n; ++i) {
            a = (a + raw[pos + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        z.insert(z.end(), raw.begin() + (std::ptrdiff_t)pos, raw.begin() + (std::ptrdiff_t)(pos + n));
        pos += n;
        if (last) break;
    }
    put32be(z, (b << 16) | a);

    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32be(ihdr, (uint32_t)w); put32be(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", z);
    chunk(png, "IEND", {});
    return dump(fn, png);
}

bool writeBMP(const std::string& fn, int w, int h, const uint8_t* rgb)
{
    if (w < 1 || h < 1 || !rgb) return false;
    const uint32_t stride = ((uint32_t)w * 3 + 3) & ~3u;
    const uint32_t size = 54 + stride * (uint32_t)h;
    std::vector<uint8_t> out(size, 0);
    auto le32 = [&](size_t o, uint32_t v) { out[o] = v; out[o + 1] = v >> 8; out[o + 2] = v >> 16; out[o + 3] = v >> 24; };
    out[0] = 'B'; out[1] = 'M';
    le32(2, size); le32(10, 54); le32(14, 40); le32(18, (uint32_t)w); le32(22, (uint32_t)h);
    out[26] = 1; out[28] = 24; le32(34, stride * (uint32_t)h);
    for (int y = 0; y < h; ++y) {                       // BMP rows are bottom-up
        const uint8_t* src = rgb + (size_t)(h - 1 - y) * w * 3;
        uint8_t* dst = out.data() + 54 + (size_t)y * stride;
        for (int x = 0; x < w; ++x) { dst[x * 3 + 0] = src[x * 3 + 2]; dst[x * 3 + 1] = src[x * 3 + 1]; dst[x * 3 + 2] = src[x * 3 + 0]; }
    }
    return dump(fn, out);
}

bool writePPM(const std::string& fn, int w, int h, const uint8_t* rgb)
{
    if (w < 1 || h < 1 || !rgb) return false;
    char hdr[64];
    const int n = std::snprintf(hdr, sizeof hdr, "P6\n%d %d\n255\n", w, h);
    std::vector<uint8_t> out(hdr, hdr + n);
    out.insert(out.end(), rgb, rgb + (size_t)w * h * 3);
    return dump(fn, out);
}

}  // namespace vr
