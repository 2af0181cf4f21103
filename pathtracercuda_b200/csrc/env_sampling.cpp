#include "env_sampling.h"
#include <cmath>
#include <cstring>

namespace ptb
{
void buildEnvDistribution(uint32_t width, uint32_t height, bool isHdr, const void *texels, EnvDistribution &out)
{
	const uint32_t bw = (width + 511u) / 512u, bh = (height + 255u) / 256u; // texels per cell
	const uint32_t cols = (width + bw - 1u) / bw, rows = (height + bh - 1u) / bh;
	const size_t n = size_t(cols) * rows;
	out.cols = cols;
	out.rows = rows;
	std::vector<double> w(n, 0.0);
	const float *f = static_cast<const float *>(texels);
	const uint8_t *b = static_cast<const uint8_t *>(texels);
	const double pi = 3.14159265358979323846;
	for (uint32_t y = 0; y < height; ++y)
	{
		const double sinTheta = sin(pi * (double(y) + 0.5) / double(height));
		for (uint32_t x = 0; x < width; ++x)
		{
			const size_t t = (size_t(y) * width + x) * 4;
			double r, g, bl;
			if (isHdr) { r = f[t]; g = f[t + 1]; bl = f[t + 2]; }
			else { r = b[t] / 255.0; g = b[t + 1] / 255.0; bl = b[t + 2] / 255.0; }
			double lum = 0.2126 * r + 0.7152 * g + 0.0722 * bl;
			if (!(lum > 0.0) || !(lum < 1e30)) lum = 0.0; // negative, NaN, infinite texels carry no weight
			w[size_t(y / bh) * cols + x / bw] += lum * sinTheta;
		}
	}
	double total = 0.0;
	for (size_t i = 0; i < n; ++i) total += w[i];
	if (!(total > 0.0)) { for (size_t i = 0; i < n; ++i) w[i] = 1.0; total = double(n); } // a black sky: uniform (never used: it contributes nothing)
	// every cell stays reachable: the bilinear filter spreads a bright texel into its dark neighbours
	const double floorW = 1e-4 * total / double(n);
	double total2 = 0.0;
	for (size_t i = 0; i < n; ++i) { w[i] += floorW; total2 += w[i]; }
	out.density.resize(n);
	std::vector<double> scaled(n);
	for (size_t i = 0; i < n; ++i)
	{
		const double p = w[i] / total2;
		out.density[i] = float(p * double(n) / (2.0 * pi * pi));
		scaled[i] = p * double(n);
	}
	// Vose's alias method; the work lists are filled in index order and used as stacks
	out.alias.assign(2 * n, 0u);
	std::vector<uint32_t> small, large;
	small.reserve(n);
	large.reserve(n);
	for (size_t i = 0; i < n; ++i) (scaled[i] < 1.0 ? small : large).push_back(uint32_t(i));
	auto put = [&](uint32_t cell, double q, uint32_t other)
	{
		const float qf = float(q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q));
		memcpy(&out.alias[2 * size_t(cell)], &qf, 4);
		out.alias[2 * size_t(cell) + 1] = other;
	};
	while (!small.empty() && !large.empty())
	{
		const uint32_t s = small.back(), l = large.back();
		small.pop_back();
		put(s, scaled[s], l);
		scaled[l] = (scaled[l] + scaled[s]) - 1.0;
		if (scaled[l] < 1.0) { large.pop_back(); small.push_back(l); }
	}
	for (uint32_t i : large) put(i, 1.0, i);
	for (uint32_t i : small) put(i, 1.0, i); // only through rounding
}
} // namespace ptb
