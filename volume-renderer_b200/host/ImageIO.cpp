// ImageIO.cpp -- see ImageIO.h.
#include "ImageIO.h"

#include <cstdio>
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
