// ImageIO.h -- the save step after the path (RendererCore::saveImage, RendererCore.cpp:165-182):
// RGB8 rows, already flipped to top-down, written as .png (stored/uncompressed deflate),
// .jpg (baseline, 4:4:4, quality 100 as the reference asks stb for), .bmp or .ppm.  The reference delegates to the vendored stb_image_write; this is a small
// self-contained writer with the same file-level results (image content, orientation).
#pragma once

#include <cstdint>
#include <string>

namespace vr {

bool writePNG(const std::string& fn, int w, int h, const uint8_t* rgb_top_down);
bool writeJPG(const std::string& fn, int w, int h, const uint8_t* rgb_top_down);
bool writeBMP(const std::string& fn, int w, int h, const uint8_t* rgb_top_down);
bool writePPM(const std::string& fn, int w, int h, const uint8_t* rgb_top_down);

}  // namespace vr
