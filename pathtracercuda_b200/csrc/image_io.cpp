#include "image_io.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <zlib.h>

namespace ptb
{
namespace
{
bool readFile(const char *path, std::vector<uint8_t> &out)
{
	FILE *f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (n < 0) { fclose(f); return false; }
	out.resize(size_t(n));
	size_t got = n ? fread(out.data(), 1, size_t(n), f) : 0;
	fclose(f);
	return got == size_t(n);
}
uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }
int paeth(int a, int b, int c)
{
	int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
} // namespace

bool decodePng(const uint8_t *d, size_t n, Image &out, std::string &err)
{
	static const uint8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
	if (n < 8 || memcmp(d, sig, 8)) { err = "not a PNG"; return false; }
	size_t p = 8;
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<uint8_t> idat, palette, trns;
	bool gotHdr = false;
	while (p + 8 <= n)
	{
		uint32_t len = be32(d + p);
		const uint8_t *type = d + p + 4;
		if (p + 12 + size_t(len) > n) { err = "truncated PNG chunk"; return false; }
		const uint8_t *body = d + p + 8;
		if (!memcmp(type, "IHDR", 4))
		{
			if (len < 13) { err = "bad IHDR"; return false; }
			w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
			gotHdr = true;
		}
		else if (!memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
		else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
		else if (!memcmp(type, "IEND", 4)) break;
		p += 12 + size_t(len);
	}
	if (!gotHdr || w == 0 || h == 0) { err = "PNG without IHDR"; return false; }
	if (interlace) { err = "interlaced PNG not supported"; return false; }
	int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!channels || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "unsupported PNG format"; return false; }
	// the colour-type / bit-depth pairs the PNG specification allows: sub-byte depths only for greyscale and palette images,
	// 16 bits not for palettes
	if ((depth < 8 && !(ctype == 0 || ctype == 3)) || (depth == 16 && ctype == 3)) { err = "invalid PNG colour type / bit depth combination"; return false; }
	// a corrupt or hostile IHDR must not turn into a giant allocation (stb_image's own limit is 2^24 per side)
	if (w > (1u << 24) || h > (1u << 24) || uint64_t(w) * h > (1ull << 28)) { err = "PNG dimensions too large"; return false; }
	if (ctype == 3 && palette.empty()) { err = "paletted PNG without PLTE"; return false; }
	const size_t bpp = std::max<size_t>(1, size_t(channels) * depth / 8);
	const size_t stride = (size_t(w) * channels * depth + 7) / 8;
	std::vector<uint8_t> raw((stride + 1) * h);
	uLongf rawLen = uLongf(raw.size());
	int zr = uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size()));
	if (zr != Z_OK || rawLen != raw.size()) { err = "PNG inflate failed"; return false; }
	// unfilter in place
	std::vector<uint8_t> zero(stride, 0);
	for (uint32_t y = 0; y < h; ++y)
	{
		uint8_t *row = raw.data() + size_t(y) * (stride + 1);
		const int f = row[0];
		uint8_t *cur = row + 1;
		const uint8_t *up = y ? (raw.data() + size_t(y - 1) * (stride + 1) + 1) : zero.data();
		for (size_t x = 0; x < stride; ++x)
		{
			const int a = x >= bpp ? cur[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
			int v = cur[x];
			switch (f)
			{
			case 0: break;
			case 1: v += a; break;
			case 2: v += b; break;
			case 3: v += (a + b) >> 1; break;
			case 4: v += paeth(a, b, c); break;
			default: err = "bad PNG filter"; return false;
			}
			cur[x] = uint8_t(v);
		}
	}
	out.width = w; out.height = h; out.isHdr = false;
	out.hdr.clear();
	out.ldr.assign(size_t(w) * h * 4, 255);
	for (uint32_t y = 0; y < h; ++y)
	{
		const uint8_t *row = raw.data() + size_t(y) * (stride + 1) + 1;
		uint8_t *dst = out.ldr.data() + size_t(y) * w * 4;
		for (uint32_t x = 0; x < w; ++x)
		{
			uint8_t s[4] = { 0, 0, 0, 255 };
			for (int c = 0; c < channels; ++c)
			{
				if (depth == 8) s[c] = row[size_t(x) * channels + c];
				else if (depth == 16) s[c] = row[(size_t(x) * channels + c) * 2]; // high byte, as stb's 16->8 conversion
				else
				{
					const size_t bit = size_t(x) * depth;
					const int v = (row[bit >> 3] >> (8 - depth - int(bit & 7))) & ((1 << depth) - 1);
					s[c] = ctype == 3 ? uint8_t(v) : uint8_t(v * (255 / ((1 << depth) - 1)));
				}
			}
			if (ctype == 3)
			{
				const size_t idx = s[0];
				dst[x * 4 + 0] = idx * 3 + 2 < palette.size() ? palette[idx * 3] : 0;
				dst[x * 4 + 1] = idx * 3 + 2 < palette.size() ? palette[idx * 3 + 1] : 0;
				dst[x * 4 + 2] = idx * 3 + 2 < palette.size() ? palette[idx * 3 + 2] : 0;
				dst[x * 4 + 3] = idx < trns.size() ? trns[idx] : 255;
			}
			else if (ctype == 0) { dst[x * 4] = dst[x * 4 + 1] = dst[x * 4 + 2] = s[0]; }
			else if (ctype == 4) { dst[x * 4] = dst[x * 4 + 1] = dst[x * 4 + 2] = s[0]; dst[x * 4 + 3] = s[1]; }
			else if (ctype == 2) { dst[x * 4] = s[0]; dst[x * 4 + 1] = s[1]; dst[x * 4 + 2] = s[2]; }
			else { dst[x * 4] = s[0]; dst[x * 4 + 1] = s[1]; dst[x * 4 + 2] = s[2]; dst[x * 4 + 3] = s[3]; }
		}
	}
	return true;
}

bool decodeHdr(const uint8_t *d, size_t n, Image &out, std::string &err)
{
	size_t p = 0;
	auto line = [&](std::string &s) -> bool
	{
		s.clear();
		if (p >= n) return false;
		while (p < n && d[p] != '\n') s += char(d[p++]);
		if (p < n) ++p;
		return true;
	};
	std::string s;
	if (!line(s) || (s != "#?RADIANCE" && s != "#?RGBE")) { err = "not a Radiance HDR file"; return false; }
	bool format = false;
	while (line(s))
	{
		if (s.empty()) break;
		if (s == "FORMAT=32-bit_rle_rgbe") format = true;
	}
	if (!format) { err = "unsupported HDR format"; return false; }
	if (!line(s)) { err = "HDR without resolution line"; return false; }
	int h = 0, w = 0;
	if (sscanf(s.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { err = "unsupported HDR data layout"; return false; }
	out.width = uint32_t(w); out.height = uint32_t(h); out.isHdr = true;
	out.ldr.clear();
	out.hdr.assign(size_t(w) * h * 4, 1.0f);
	std::vector<uint8_t> rgbe(size_t(w) * h * 4);
	bool flat = (w < 8 || w >= 32768);
	size_t y0 = 0;
	if (!flat)
	{
		std::vector<uint8_t> scan(size_t(w) * 4);
		for (int y = 0; y < h; ++y)
		{
			if (p + 4 > n) { err = "truncated HDR"; return false; }
			const int c1 = d[p], c2 = d[p + 1], len = d[p + 2];
			if (c1 != 2 || c2 != 2 || (len & 0x80))
			{
				if (y != 0) { err = "corrupt HDR scanline"; return false; }
				flat = true; // not run-length encoded: the whole image is flat RGBE starting here
				break;
			}
			if (((len << 8) | d[p + 3]) != w) { err = "invalid decoded scanline length"; return false; }
			p += 4;
			for (int c = 0; c < 4; ++c)
			{
				int x = 0;
				while (x < w)
				{
					if (p >= n) { err = "truncated HDR"; return false; }
					int count = d[p++];
					if (count > 128)
					{
						count -= 128;
						if (p >= n || x + count > w) { err = "corrupt HDR run"; return false; }
						const uint8_t v = d[p++];
						for (int k = 0; k < count; ++k) scan[size_t(x++) * 4 + c] = v;
					}
					else
					{
						if (count == 0 || p + size_t(count) > n || x + count > w) { err = "corrupt HDR run"; return false; }
						for (int k = 0; k < count; ++k) scan[size_t(x++) * 4 + c] = d[p++];
					}
				}
			}
			memcpy(rgbe.data() + size_t(y) * w * 4, scan.data(), scan.size());
			y0 = size_t(y) + 1;
		}
	}
	if (flat)
	{
		const size_t need = (size_t(h) - y0) * w * 4;
		if (p + need > n) { err = "truncated HDR"; return false; }
		memcpy(rgbe.data() + y0 * w * 4, d + p, need);
	}
	for (size_t i = 0; i < size_t(w) * h; ++i)
	{
		const uint8_t *in = rgbe.data() + i * 4;
		float *o = out.hdr.data() + i * 4;
		if (in[3] != 0)
		{
			const float f1 = ldexpf(1.0f, int(in[3]) - (128 + 8));
			o[0] = in[0] * f1; o[1] = in[1] * f1; o[2] = in[2] * f1;
		}
		else { o[0] = o[1] = o[2] = 0.0f; }
		o[3] = 1.0f;
	}
	return true;
}

bool readImage(const char *path, Image &out, std::string &err)
{
	std::vector<uint8_t> data;
	if (!readFile(path, data)) { err = std::string("cannot read ") + path; return false; }
	const bool hdr = (data.size() >= 11 && !memcmp(data.data(), "#?RADIANCE\n", 11)) || (data.size() >= 7 && !memcmp(data.data(), "#?RGBE\n", 7));
	if (hdr) return decodeHdr(data.data(), data.size(), out, err);
	static const uint8_t pngSig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
	if (data.size() < 8 || memcmp(data.data(), pngSig, 8) != 0)
	{
		// the reference decodes textures with stb_image (also JPEG, BMP, TGA, PSD, GIF, PIC, PNM); here: PNG and Radiance HDR.
		// Say so instead of silently rendering untextured.
		const char *what = data.size() >= 3 && data[0] == 0xff && data[1] == 0xd8 ? "JPEG" : data.size() >= 2 && data[0] == 'B' && data[1] == 'M' ? "BMP"
		                   : data.size() >= 4 && !memcmp(data.data(), "GIF8", 4) ? "GIF" : data.size() >= 4 && !memcmp(data.data(), "8BPS", 4) ? "PSD" : "unknown";
		err = std::string("unsupported image format (") + what + "; supported: PNG, Radiance HDR): " + path;
		fprintf(stderr, "pt_b200: %s\n", err.c_str());
		return false;
	}
	const bool ok = decodePng(data.data(), data.size(), out, err);
	if (!ok && err == "interlaced PNG not supported") fprintf(stderr, "pt_b200: %s: %s\n", err.c_str(), path);
	return ok;
}

// ---- writers ------------------------------------------------------------------------------------------------
namespace
{
void putChunk(std::vector<uint8_t> &o, const char *type, const uint8_t *body, size_t len)
{
	uint8_t hdr[8] = { uint8_t(len >> 24), uint8_t(len >> 16), uint8_t(len >> 8), uint8_t(len), uint8_t(type[0]), uint8_t(type[1]), uint8_t(type[2]), uint8_t(type[3]) };
	o.insert(o.end(), hdr, hdr + 8);
	if (len) o.insert(o.end(), body, body + len);
	uLong crc = crc32(0L, hdr + 4, 4);
	if (len) crc = crc32(crc, body, uInt(len));
	uint8_t c[4] = { uint8_t(crc >> 24), uint8_t(crc >> 16), uint8_t(crc >> 8), uint8_t(crc) };
	o.insert(o.end(), c, c + 4);
}
} // namespace

bool writePng(const char *path, uint32_t w, uint32_t h, const uint8_t *rgba, std::string &err)
{
	const size_t stride = size_t(w) * 4;
	std::vector<uint8_t> filtered((stride + 1) * h), cand(stride);
	std::vector<uint8_t> zero(stride, 0);
	for (uint32_t y = 0; y < h; ++y)
	{
		const uint8_t *cur = rgba + size_t(h - 1 - y) * stride; // flip: file row 0 = top of the view
		const uint8_t *up = y ? rgba + size_t(h - y) * stride : zero.data();
		int bestF = 0;
		long bestSum = -1;
		uint8_t *dst = filtered.data() + size_t(y) * (stride + 1);
		for (int f = 0; f < 5; ++f) // minimum-sum-of-absolute-differences heuristic
		{
			long sum = 0;
			for (size_t x = 0; x < stride; ++x)
			{
				const int a = x >= 4 ? cur[x - 4] : 0, b = up[x], c = x >= 4 ? up[x - 4] : 0;
				int v = cur[x];
				if (f == 1) v -= a; else if (f == 2) v -= b; else if (f == 3) v -= (a + b) >> 1; else if (f == 4) v -= paeth(a, b, c);
				cand[x] = uint8_t(v);
				sum += abs(int(int8_t(cand[x])));
			}
			if (bestSum < 0 || sum < bestSum) { bestSum = sum; bestF = f; dst[0] = uint8_t(f); memcpy(dst + 1, cand.data(), stride); }
		}
		(void)bestF;
	}
	uLongf clen = compressBound(uLong(filtered.size()));
	std::vector<uint8_t> comp(clen);
	if (compress2(comp.data(), &clen, filtered.data(), uLong(filtered.size()), 6) != Z_OK) { err = "deflate failed"; return false; }
	std::vector<uint8_t> o = { 137, 80, 78, 71, 13, 10, 26, 10 };
	uint8_t ihdr[13] = { uint8_t(w >> 24), uint8_t(w >> 16), uint8_t(w >> 8), uint8_t(w), uint8_t(h >> 24), uint8_t(h >> 16), uint8_t(h >> 8), uint8_t(h), 8, 6, 0, 0, 0 };
	putChunk(o, "IHDR", ihdr, 13);
	putChunk(o, "IDAT", comp.data(), clen);
	putChunk(o, "IEND", nullptr, 0);
	FILE *f = fopen(path, "wb");
	if (!f) { err = std::string("cannot write ") + path; return false; }
	const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
	fclose(f);
	if (!ok) err = "short write";
	return ok;
}

bool writeHdr(const char *path, uint32_t w, uint32_t h, const float *rgba, std::string &err)
{
	FILE *f = fopen(path, "wb");
	if (!f) { err = std::string("cannot write ") + path; return false; }
	// header text as written by the reference's encoder
	fprintf(f, "#?RADIANCE\n# Written by stb_image_write.h\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y %u +X %u\n", h, w);
	std::vector<uint8_t> rgbe(size_t(w) * 4), out;
	for (uint32_t y = 0; y < h; ++y)
	{
		const float *row = rgba + size_t(h - 1 - y) * w * 4; // flip on write
		for (uint32_t x = 0; x < w; ++x)
		{
			const float r = row[x * 4], g = row[x * 4 + 1], b = row[x * 4 + 2];
			const float m = r > g ? (r > b ? r : b) : (g > b ? g : b);
			uint8_t *e = rgbe.data() + size_t(x) * 4;
			if (m < 1e-32f) { e[0] = e[1] = e[2] = e[3] = 0; }
			else
			{
				int ex;
				const float norm = float(frexp(m, &ex)) * 256.0f / m;
				e[0] = (unsigned char)(r * norm); e[1] = (unsigned char)(g * norm); e[2] = (unsigned char)(b * norm); e[3] = (unsigned char)(ex + 128);
			}
		}
		out.clear();
		if (w < 8 || w >= 32768) out.insert(out.end(), rgbe.begin(), rgbe.end());
		else
		{
			out.push_back(2); out.push_back(2); out.push_back(uint8_t(w >> 8)); out.push_back(uint8_t(w & 0xff));
			for (int c = 0; c < 4; ++c)
			{
				uint32_t x = 0;
				while (x < w)
				{
					// find the next run of >= 3 equal bytes
					uint32_t r = x;
					while (r + 2 < w && !(rgbe[r * 4 + c] == rgbe[(r + 1) * 4 + c] && rgbe[r * 4 + c] == rgbe[(r + 2) * 4 + c])) ++r;
					if (r + 2 >= w) r = w;
					while (x < r) // literal bytes before the run
					{
						uint32_t len = r - x; if (len > 128) len = 128;
						out.push_back(uint8_t(len));
						for (uint32_t k = 0; k < len; ++k) out.push_back(rgbe[(x + k) * 4 + c]);
						x += len;
					}
					if (r + 2 < w)
					{
						uint32_t e = r + 2;
						while (e < w && rgbe[e * 4 + c] == rgbe[r * 4 + c]) ++e;
						while (x < e)
						{
							uint32_t len = e - x; if (len > 127) len = 127;
							out.push_back(uint8_t(len + 128));
							out.push_back(rgbe[r * 4 + c]);
							x += len;
						}
					}
				}
			}
		}
		if (fwrite(out.data(), 1, out.size(), f) != out.size()) { fclose(f); err = "short write"; return false; }
	}
	fclose(f);
	return true;
}
} // namespace ptb
