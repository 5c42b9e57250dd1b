// VolumeIO.cpp -- see VolumeIO.h.
#include "VolumeIO.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
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

// undo the channel de-interleave: the stream stores all bytes of channel 0, then channel 1 ...
void weaveChannels(std::vector<uint8_t>& data, uint32_t channels, uint64_t block)
{
    if (channels <= 1) return;
    const uint64_t total = data.size();
    const uint64_t span = block ? (uint64_t)channels * block : total;
    std::vector<uint8_t> tmp;
    for (uint64_t base = 0; base < total; base += span) {
        const uint64_t len = (total - base < span) ? total - base : span;
        tmp.assign(data.begin() + (std::ptrdiff_t)base, data.begin() + (std::ptrdiff_t)(base + len));
        uint8_t* d = data.data() + base;
        if (channels == 2 && (len & 1) == 0) {
            // 16-bit volumes: the two byte planes of the block, zipped
            const uint8_t* lo = tmp.data();
            const uint8_t* hi = tmp.data() + len / 2;
            for (uint64_t j = 0; j < len / 2; ++j) { d[2 * j] = lo[j]; d[2 * j + 1] = hi[j]; }
        } else {
            uint64_t src = 0;
            for (uint32_t c = 0; c < channels; ++c)
                for (uint64_t j = c; j < len; j += channels) d[j] = tmp[src++];
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
    uint64_t cap = size * 2 + 4096;
    out.resize(cap);
    uint8_t* o = out.data();
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
            cap = cap * 2 + run;
            out.resize(cap);
            o = out.data();
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
                const unsigned per_fill = 56 / width;             // fields guaranteed after one fill()
                while (k < run) {
                    bits.fill();
                    const uint32_t m = (run - k < per_fill) ? run - k : per_fill;
                    for (uint32_t j = 0; j < m; ++j, ++h) {
                        const int delta = (int)bits.takeUnchecked(width) - bias + (int)h[0] - (int)h[-1];
                        value = (value + delta) & 0xff;           // wrap into 0..255
                        o[n++] = (uint8_t)value;
                    }
                    k += m;
                }
            }
        }
    }
    // a stream without its end marker decodes zero padding until a zero count turns up: detect it here
    if (bits.exhausted()) { error = "DDS stream: missing end-of-stream marker"; return false; }
    out.resize(n);
    weaveChannels(out, channels, block);
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
    out.payload.assign(reinterpret_cast<const uint8_t*>(p), reinterpret_cast<const uint8_t*>(p) + vol);
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
