// VolumeIO.cpp -- see VolumeIO.h.
#include "VolumeIO.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <new>
#include <sstream>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace vr {

// ------------------------------------------------------------------ .raw + .raw.inf

bool readRawInf(const std::string& raw_fn, RawInf& out, bool& exists, std::string& title, std::string& msg)
{
    std::ifstream f(raw_fn + ".inf");
    exists = (bool)f;
    if (!exists) return false;
    out = RawInf();
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        if (line == "#dimensions") {
            std::getline(f, line);
            if (line.empty()) {
                msg = "Dimensions for Volume Data not provided in \"raw.inf\" file.";
                title = "Invalid .raw.inf file!";
                return false;
            }
            std::stringstream ss(line);
            ss >> out.dims[0]; ss >> out.dims[1]; ss >> out.dims[2];
        } else if (line == "#voxel-spacing") {
            std::getline(f, line);
            if (line.empty()) {
                msg = "Aspect Ratio for Volume Data not provided in \"raw.inf\" file.";
                title = "Invalid .raw.inf file!";
                return false;
            }
            std::stringstream ss(line);
            ss >> out.spacing[0]; ss >> out.spacing[1]; ss >> out.spacing[2];
        }
    }
    if (out.dims[0] == 0 && out.dims[1] == 0 && out.dims[2] == 0) {
        msg = "Dimensions for Volume Data not provided in \"raw.inf\" file. Make sure the header is \"#dimesnsions\"";
        title = "Invalid .raw.inf file!";
        return false;
    }
    if (out.spacing[0] == 0 && out.spacing[1] == 0 && out.spacing[2] == 0) {
        msg = "Aspect Ratio for Volume Data not provided in \"raw.inf\" file. Make sure the header is \"#voxel-spacing\"";
        title = "Invalid .raw.inf file!";
        return false;
    }
    return true;
}

bool writeRawInf(const std::string& raw_fn, const RawInf& inf)
{
    std::ofstream o(raw_fn + ".inf");
    if (!o) return false;
    o << "#dimensions\n" << inf.dims[0] << " " << inf.dims[1] << " " << inf.dims[2] << "\n\n"
      << "#voxel-spacing\n" << inf.spacing[0] << " " << inf.spacing[1] << " " << inf.spacing[2] << std::endl;
    return (bool)o;
}

bool readRawPayload(const std::string& raw_fn, uint64_t n, std::vector<uint8_t>& out)
{
    std::ifstream f(raw_fn, std::ios::binary);
    if (!f) return false;
    out.assign(n, 0);
    f.read(reinterpret_cast<char*>(out.data()), (std::streamsize)n);
    return true;
}

MappedFile::~MappedFile() { close(); }

void MappedFile::close()
{
    if (data_ && size_) ::munmap(const_cast<uint8_t*>(data_), (size_t)size_);
    data_ = nullptr; size_ = 0;
}

bool MappedFile::open(const std::string& fn, std::string& error)
{
    close();
    const int fd = ::open(fn.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) { error = "cannot open " + fn + ": " + std::strerror(errno); return false; }
    struct stat st;
    if (::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); error = fn + " is not a regular file"; return false; }
    if (st.st_size == 0) { ::close(fd); return true; }          // empty file: valid, nothing to map
    void* p = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);                                                 // the mapping keeps the file alive
    if (p == MAP_FAILED) { error = "mmap of " + fn + " failed: " + std::strerror(errno); return false; }
    ::madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
    data_ = static_cast<const uint8_t*>(p); size_ = (uint64_t)st.st_size;
    return true;
}

bool checkedVolumeBytes(const uint64_t dims[3], uint64_t bytes_per_voxel, uint64_t& bytes)
{
    bytes = 0;
    if (bytes_per_voxel < 1 || bytes_per_voxel > 16) return false;
    for (int i = 0; i < 3; ++i) if (dims[i] < 1 || dims[i] > kMaxVolumeDim) return false;
    // 16384^3 * 16 = 2^46: cannot overflow once every factor is range checked
    bytes = dims[0] * dims[1] * dims[2] * bytes_per_voxel;
    return true;
}

// ------------------------------------------------------------------ DDS bit stream

namespace {

// MSB-first bit source over a byte buffer with a 64-bit accumulator (left aligned).  Reads past
// the end yield zero bits (the reference pads its cache with a zero word).
class BitSource {
public:
    BitSource(const uint8_t* p, uint64_t n) : p_(p), n_(n) {}
    // nbits <= 32
    inline uint32_t take(unsigned nbits)
    {
        if (nbits == 0) return 0;
        if (avail_ < nbits) refill();
        const uint32_t v = (uint32_t)(acc_ >> (64 - nbits));
        acc_ <<= nbits;
        avail_ -= nbits;
        return v;
    }
    // true once more than four bytes of padding have been consumed: the stream has no end marker
    bool exhausted() const { return pos_ * 8 - avail_ > (n_ + 4) * 8; }
    // make at least 56 bits available, then hand out `nbits`-wide fields without further checks
    inline void fill() { if (avail_ < 56) refill(); }
    inline unsigned available() const { return avail_; }
    inline uint32_t takeUnchecked(unsigned nbits)      // 1 <= nbits <= 8, caller guarantees availability
    {
        const uint32_t v = (uint32_t)(acc_ >> (64 - nbits));
        acc_ <<= nbits;
        avail_ -= nbits;
        return v;
    }
    // the accumulator for a hot loop that keeps it in a register: after fill(), the caller shifts `nbits` out of the
    // returned word itself and reports what it consumed (the decoder's output pointer is a byte pointer and may alias
    // this object as far as the compiler knows, so member updates per field would go through memory)
    inline uint64_t peek() const { return acc_; }
    inline void consumed(uint64_t acc_after, unsigned nbits) { acc_ = acc_after; avail_ -= nbits; }
private:
    inline void refill()
    {
        if (pos_ + 8 <= n_) {
            // eight bytes at once, big endian; keep what fits behind the bits still pending
            uint64_t w;
            std::memcpy(&w, p_ + pos_, 8);
            w = __builtin_bswap64(w);
            const unsigned take_bytes = (64 - avail_) >> 3;
            if (take_bytes == 8) acc_ = w;                       // avail_ == 0: a 64-bit shift would be undefined
            else acc_ |= (w >> avail_) & ~((1ull << (64 - avail_ - 8 * take_bytes)) - 1ull);
            pos_ += take_bytes;
            avail_ += 8 * take_bytes;
            return;
        }
        while (avail_ <= 56) {
            const uint64_t b = pos_ < n_ ? p_[pos_] : 0;
            ++pos_;
            acc_ |= b << (56 - avail_);
            avail_ += 8;
        }
    }
    const uint8_t* p_;
    uint64_t n_, pos_ = 0;
    uint64_t acc_ = 0;
    unsigned avail_ = 0;
};

// `count` predicted values of constant field width W (2..8): value += field - bias + (h[0] - h[-1]), wrapped to a byte.
// The width is a template parameter so that every shift has an immediate count and the per-fill group unrolls; the
// accumulator lives in a register (see BitSource::peek).  Returns the running value.
template <unsigned W>
inline int decodePredictedRun(BitSource& bits, const uint8_t* h, uint8_t* w, uint32_t count, int value)
{
    constexpr unsigned PER_FILL = 56 / W;             // fields guaranteed after one fill()
    constexpr int BIAS = (int)((1u << W) >> 1);
    while (count >= PER_FILL) {
        bits.fill();
        uint64_t acc = bits.peek();
#pragma GCC unroll 28
        for (unsigned j = 0; j < PER_FILL; ++j) {
            const int delta = (int)(uint32_t)(acc >> (64 - W)) - BIAS + (int)h[j] - (int)h[(std::ptrdiff_t)j - 1];
            acc <<= W;
            value = (value + delta) & 0xff;
            w[j] = (uint8_t)value;
        }
        bits.consumed(acc, PER_FILL * W);
        h += PER_FILL; w += PER_FILL; count -= PER_FILL;
    }
    if (count) {
        bits.fill();
        uint64_t acc = bits.peek();
        for (uint32_t j = 0; j < count; ++j) {
            const int delta = (int)(uint32_t)(acc >> (64 - W)) - BIAS + (int)h[j] - (int)h[(std::ptrdiff_t)j - 1];
            acc <<= W;
            value = (value + delta) & 0xff;
            w[j] = (uint8_t)value;
        }
        bits.consumed(acc, count * W);
    }
    return value;
}

// undo the channel de-interleave: the stream stores all bytes of channel 0, then channel 1 ... (per `block` * channels
// bytes when the file is block-interleaved, else over the whole payload).  Reads the planar bytes, writes the woven ones:
// one pass, no temporary; 16-bit volumes (two byte planes) are zipped 16 bytes at a time.
void weaveChannels(const uint8_t* planar, uint64_t total, uint32_t channels, uint64_t block, uint8_t* woven)
{
    const uint64_t span = block ? (uint64_t)channels * block : total;
    for (uint64_t base = 0; base < total; base += span) {
        const uint64_t len = (total - base < span) ? total - base : span;
        const uint8_t* src = planar + base;
        uint8_t* d = woven + base;
        if (channels == 2 && (len & 1) == 0) {
            const uint8_t* lo = src;
            const uint8_t* hi = src + len / 2;
            const uint64_t half = len / 2;
            uint64_t j = 0;
#if defined(__SSE2__)
            for (; j + 16 <= half; j += 16) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(lo + j));
                const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(hi + j));
                _mm_storeu_si128(reinterpret_cast<__m128i*>(d + 2 * j), _mm_unpacklo_epi8(a, b));
                _mm_storeu_si128(reinterpret_cast<__m128i*>(d + 2 * j + 16), _mm_unpackhi_epi8(a, b));
            }
#endif
            for (; j < half; ++j) { d[2 * j] = lo[j]; d[2 * j + 1] = hi[j]; }
        } else {
            uint64_t s = 0;
            for (uint32_t c = 0; c < channels; ++c)
                for (uint64_t j = c; j < len; j += channels) d[j] = src[s++];
        }
        if (span == total) break;
    }
}

}  // namespace

// The bit stream is sequential by construction (every value is a delta on the previous one), so the
// decoder is one tight loop: runs of `run` values share a bit width; after the first `row` values the
// delta is additionally predicted from the previous row (ddsbase.cpp:394-452 of the reference).
bool ddsDecode(const uint8_t* chunk, uint64_t size, uint64_t block, std::vector<uint8_t>& out, std::string& error)
{
    BitSource bits(chunk, size);
    const uint32_t channels = bits.take(2) + 1;     // "skip"
    const uint32_t row = bits.take(16) + 1;         // "strip": predictor distance
    // planar bytes first (channel 0, then channel 1 ...), in a buffer that is NOT zero-filled first (the decoder writes
    // every byte it later reads; a std::vector would touch all of it twice)
    uint64_t cap = size * 2 + 4096;
    std::unique_ptr<uint8_t[]> planar(new (std::nothrow) uint8_t[cap]);
    if (!planar) { error = "DDS stream: out of memory"; return false; }
    uint8_t* o = planar.get();
    uint64_t n = 0;
    int value = 0;
    for (;;) {
        // run header: 7-bit count (0 = end of stream) + 3-bit width code.  Both fields in one read; when the
        // count is zero the three extra bits belong to nothing (the stream ends there), so over-reading is harmless.
        const uint32_t hdr = bits.take(10);
        const uint32_t run = hdr >> 3;
        if (run == 0) break;
        const uint32_t code = hdr & 7u;
        const unsigned width = code ? code + 1 : 0;
        const int bias = (int)((1u << width) >> 1);
        if (n + run > cap) {
            const uint64_t grown = cap * 2 + run;
            std::unique_ptr<uint8_t[]> bigger(new (std::nothrow) uint8_t[grown]);
            if (!bigger) { error = "DDS stream: out of memory"; return false; }
            std::memcpy(bigger.get(), planar.get(), n);
            planar.swap(bigger);
            cap = grown;
            o = planar.get();
        }
        uint32_t k = 0;
        // values that have no predictor yet (or never: row == 1)
        for (; k < run && (row == 1 || n <= row); ++k) {
            value = (value + (int)bits.take(width) - bias) & 0xff;
            o[n++] = (uint8_t)value;
        }
        if (k < run) {
            // predicted values: delta += previous row's step.  h[0] = value one row back, h[-1] = its predecessor
            const uint8_t* h = o + n - row;
            if (width == 0) {
                for (; k < run; ++k, ++h) {
                    value = (value + (int)h[0] - (int)h[-1]) & 0xff;
                    o[n++] = (uint8_t)value;
                }
            } else {
                const uint32_t count = run - k;
                uint8_t* w = o + n;
                switch (width) {
                    case 2: value = decodePredictedRun<2>(bits, h, w, count, value); break;
                    case 3: value = decodePredictedRun<3>(bits, h, w, count, value); break;
                    case 4: value = decodePredictedRun<4>(bits, h, w, count, value); break;
                    case 5: value = decodePredictedRun<5>(bits, h, w, count, value); break;
                    case 6: value = decodePredictedRun<6>(bits, h, w, count, value); break;
                    case 7: value = decodePredictedRun<7>(bits, h, w, count, value); break;
                    default: value = decodePredictedRun<8>(bits, h, w, count, value); break;
                }
                n += count;
            }
        }
    }
    // a stream without its end marker decodes zero padding until a zero count turns up: detect it here
    if (bits.exhausted()) { error = "DDS stream: missing end-of-stream marker"; return false; }
    out.clear();
    out.reserve(n + 1);                             // + 1: the PVM parser appends a terminator without reallocating
    out.resize(n);
    if (channels > 1) weaveChannels(planar.get(), n, channels, block, out.data());
    else if (n) std::memcpy(out.data(), planar.get(), n);
    return true;
}

// ------------------------------------------------------------------ PVM container

namespace {

const char* skipLine(const char* p, const char* end)
{
    while (p < end && *p != '\n') ++p;
    return p < end ? p + 1 : nullptr;
}

bool parseUInts(const char*& p, int n, uint32_t* dst)
{
    for (int i = 0; i < n; ++i) {
        char* e = nullptr;
        errno = 0;
        const long v = std::strtol(p, &e, 10);
        if (e == p || errno) return false;
        if (v < 0 || v > 0x7fffffffL) return false;
        dst[i] = (uint32_t)v;
        p = e;
    }
    return true;
}

bool parseFloats(const char*& p, int n, float* dst)
{
    for (int i = 0; i < n; ++i) {
        char* e = nullptr;
        const float v = std::strtof(p, &e);
        if (e == p) return false;
        dst[i] = v;
        p = e;
    }
    return true;
}

}  // namespace

bool pvmDecode(const uint8_t* file, uint64_t bytes, PvmVolume& out, std::string& error)
{
    static const char kV3d[] = "DDS v3d\n", kV3e[] = "DDS v3e\n";
    std::vector<uint8_t> body;
    if (bytes >= 8 && std::memcmp(file, kV3d, 8) == 0) {
        if (!ddsDecode(file + 8, bytes - 8, 0, body, error)) return false;
    } else if (bytes >= 8 && std::memcmp(file, kV3e, 8) == 0) {
        if (!ddsDecode(file + 8, bytes - 8, 1ull << 24, body, error)) return false;
    } else {
        body.assign(file, file + bytes);            // plain, uncompressed PVM
    }
    if (body.size() < 5) { error = "PVM: file too short"; return false; }
    body.push_back(0);                              // terminator for the text parsers
    const char* base = reinterpret_cast<const char*>(body.data());
    const char* end = base + body.size() - 1;
    const char* p = nullptr;

    out = PvmVolume();
    uint32_t dims[3] = {0, 0, 0};
    if (std::strncmp(base, "PVM\n", 4) == 0) {
        out.version = 1;
        p = base + 4;
        while (p && p < end && *p == '#') p = skipLine(p, end);   // comment lines
        if (!p || !parseUInts(p, 3, dims)) { error = "PVM: bad dimension line"; return false; }
    } else if (std::strncmp(base, "PVM2\n", 5) == 0 || std::strncmp(base, "PVM3\n", 5) == 0) {
        out.version = base[3] - '0';
        p = base + 5;
        if (!parseUInts(p, 3, dims) || !parseFloats(p, 3, out.scale)) { error = "PVM: bad dimension/scale lines"; return false; }
        if (!(out.scale[0] > 0.0f) || !(out.scale[1] > 0.0f) || !(out.scale[2] > 0.0f)) { error = "PVM: non-positive voxel scale"; return false; }
    } else {
        error = "PVM: missing PVM/PVM2/PVM3 header";
        return false;
    }
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) { error = "PVM: zero dimension"; return false; }
    const uint64_t dims64[3] = {dims[0], dims[1], dims[2]};
    p = skipLine(p, end);
    uint32_t comps = 0;
    if (!p || !parseUInts(p, 1, &comps) || comps < 1) { error = "PVM: bad component count"; return false; }
    p = skipLine(p, end);
    if (!p) { error = "PVM: truncated header"; return false; }

    // untrusted header: range-check every factor before multiplying (each can be up to 2^32 - 1)
    uint64_t vol = 0;
    if (!checkedVolumeBytes(dims64, comps, vol)) { error = "PVM: dimensions outside [1,16384] or more than 16 components"; return false; }
    if ((uint64_t)(end - p) < vol) { error = "PVM: payload shorter than width*height*depth*components"; return false; }
    const char* q = p + vol;
    std::string* strs[4] = {&out.description, &out.courtesy, &out.parameter, &out.comment};
    if (out.version == 3) {
        for (int i = 0; i < 4; ++i) {
            const void* z = std::memchr(q, 0, (size_t)(end - q));
            if (!z) { error = "PVM3: unterminated trailer string"; return false; }
            strs[i]->assign(q, (const char*)z);
            q = (const char*)z + 1;
        }
    }
    if (q != end) { error = "PVM: trailing bytes after payload"; return false; }
    out.width = dims[0]; out.height = dims[1]; out.depth = dims[2]; out.components = comps;
    // the payload is `vol` bytes in the middle of `body`: move them to the front in place and hand the buffer over
    // (a copy into a fresh vector costs as much as the page faults of another width*height*depth*components bytes)
    const size_t header = (size_t)(p - base);
    if (header) std::memmove(body.data(), body.data() + header, (size_t)vol);
    body.resize((size_t)vol);
    out.payload = std::move(body);
    return true;
}

bool pvmReadFile(const std::string& fn, PvmVolume& out, std::string& error)
{
    std::ifstream f(fn, std::ios::binary | std::ios::ate);
    if (!f) { error = "cannot open " + fn; return false; }
    const std::streamoff n = f.tellg();
    f.seekg(0);
    std::vector<uint8_t> buf((size_t)n);
    if (n > 0) f.read(reinterpret_cast<char*>(buf.data()), n);
    if (!f) { error = "short read on " + fn; return false; }
    return pvmDecode(buf.data(), buf.size(), out, error);
}

uint32_t ddsChecksum(const uint8_t* data, uint64_t bytes)
{
    uint32_t sum = 0, cipher = 1;
    for (uint64_t i = 0; i < bytes; ++i) {
        cipher = 271u * cipher + data[i];
        sum += cipher * data[i];
    }
    return sum;
}

}  // namespace vr
