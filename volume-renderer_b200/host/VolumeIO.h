// VolumeIO.h -- the on-disk formats either side of the loader:
//   <file>.raw + <file>.raw.inf sidecar   (RendererCore.cpp:249-341 of the reference)
//   .pvm  = optional "DDS v3d\n" / "DDS v3e\n" differential bit stream around a PVM/PVM2/PVM3
//           text header + payload (ddsbase.cpp:394-452, 550-594, 768-858 of the reference).
// Written from the format description, 64-bit sizes throughout (the reference is limited to
// 2^31 voxels / 4 GiB).  Malformed input is reported through `error`, never printed-and-ignored.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace vr {

struct RawInf {
    int dims[3] = {0, 0, 0};
    float spacing[3] = {0, 0, 0};
};

// Returns false with title/msg set exactly like the reference's popups when the sidecar is
// missing a section (RendererCore.cpp:264-301).  `exists` tells whether the file was there.
bool readRawInf(const std::string& raw_fn, RawInf& out, bool& exists, std::string& title, std::string& msg);
// Same text the reference writes (RendererCore.cpp:311-315).
bool writeRawInf(const std::string& raw_fn, const RawInf& inf);
// Reads exactly n bytes (short files are zero filled like the reference's value-initialised buffer).
bool readRawPayload(const std::string& raw_fn, uint64_t n, std::vector<uint8_t>& out);

// Read-only memory map of a payload file (the `.raw` path of SURVEY.md 8f-3): no host copy of a multi-GiB
// volume is made -- the pages stream from the page cache straight into the upload -- and every size is 64-bit
// (the reference reads through `int len`, RendererCore.cpp:327, and fails above 2^31 voxels).
class MappedFile {
public:
    MappedFile() = default;
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    ~MappedFile();
    bool open(const std::string& fn, std::string& error);
    void close();
    const uint8_t* data() const { return data_; }
    uint64_t size() const { return size_; }
private:
    const uint8_t* data_ = nullptr;
    uint64_t size_ = 0;
};

// per-dimension limit shared with the C-ABI upload (volren_b200.h: each dimension in [1,16384]) and an
// overflow-checked byte count; false when a dimension is out of range or the product does not fit in 63 bits
constexpr uint64_t kMaxVolumeDim = 16384;
bool checkedVolumeBytes(const uint64_t dims[3], uint64_t bytes_per_voxel, uint64_t& bytes);

struct PvmVolume {
    uint32_t width = 0, height = 0, depth = 0, components = 0;
    float scale[3] = {1.0f, 1.0f, 1.0f};
    int version = 1;                               // 1 = PVM, 2 = PVM2, 3 = PVM3
    std::vector<uint8_t> payload;                  // width*height*depth*components bytes
    std::string description, courtesy, parameter, comment;   // PVM3 only
};

// Decode a DDS differential stream (the bytes after the 8-byte magic).  block = 0 for v3d,
// 1<<24 for v3e.
bool ddsDecode(const uint8_t* chunk, uint64_t size, uint64_t block, std::vector<uint8_t>& out, std::string& error);
// Decode a whole .pvm file image held in memory (DDS-wrapped or plain).
bool pvmDecode(const uint8_t* file, uint64_t bytes, PvmVolume& out, std::string& error);
bool pvmReadFile(const std::string& fn, PvmVolume& out, std::string& error);
// ddsbase.cpp:872-893 digest, kept because it is the reference's own fingerprint of a payload
uint32_t ddsChecksum(const uint8_t* data, uint64_t bytes);

}  // namespace vr
