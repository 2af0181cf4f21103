// image_io.h — PNG / Radiance-HDR codecs at the drop-in surface.  The reference uses the vendored
// stb_image 2.26 / stb_image_write 1.15 (third party, not copied); these are independent implementations on zlib that
// DECODE to the same texels and ENCODE files that decode to the same pixels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ptb
{
struct Image
{
	uint32_t width = 0, height = 0;
	bool isHdr = false;
	std::vector<uint8_t> ldr; // RGBA8, top row first (stbi_load(..., 4))
	std::vector<float> hdr;   // RGBA float, top row first (stbi_loadf(..., 4))
};

// stbi_is_hdr + stbi_load / stbi_loadf (reference Pathtracer.cpp:245-251).  PNG and Radiance .hdr only.
bool readImage(const char *path, Image &out, std::string &err);
bool decodePng(const uint8_t *data, size_t size, Image &out, std::string &err);
bool decodeHdr(const uint8_t *data, size_t size, Image &out, std::string &err);

// stbi_write_png / stbi_write_hdr with stbi_flip_vertically_on_write(true) (reference main.cpp:184-194):
// `rgba` is bottom-row-first; the file's first scanline is the buffer's last row.
bool writePng(const char *path, uint32_t w, uint32_t h, const uint8_t *rgba, std::string &err);
bool writeHdr(const char *path, uint32_t w, uint32_t h, const float *rgba, std::string &err);
} // namespace ptb
