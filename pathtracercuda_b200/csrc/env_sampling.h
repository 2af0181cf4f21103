// env_sampling.h — importance sampling of the environment map (north_star item 5; option "env_is").
//
// The reference evaluates the sky only where a path misses (kernels/trace.cu:115-134) and samples directions from the BSDF
// alone.  With "env_is" every scattering vertex ALSO draws one direction from a distribution proportional to the sky's
// radiance and the two estimates are combined by multiple importance sampling (balance heuristic) - the same integral, paths of
// at most five segments, sky light only through a miss on segments 2..5 of a path plus the camera ray's own miss - so the image
// converges to the reference's, with less variance per sample wherever the sky has a sun in it.
//
// This file builds the distribution on the host: a grid of at most 512 x 256 cells over the equirectangular map, cell weight =
// sum over its texels of luminance x sin(theta), as ONE alias table over all cells (a sample is one 8-byte read on the read-only
// path, no binary search) and a per-cell density table (one 4-byte read to evaluate the pdf of any direction).
#pragma once
#include <cstdint>
#include <vector>

namespace ptb
{
struct EnvDistribution
{
	uint32_t cols = 0, rows = 0;
	std::vector<uint32_t> alias; // 2 words per cell: float bits of the acceptance threshold q, alias cell index
	std::vector<float> density;  // P(cell) * cols * rows / (2 pi^2): the solid-angle pdf of a direction in the cell is density / sin(theta)
};

// texels: width x height, RGBA float (is_hdr) or RGBA8.  Deterministic (double-precision sums in texel order, Vose's alias
// method with index-ordered work lists): the oracle (oracle/pt_oracle.c) restates it and gets the same tables bit for bit.
void buildEnvDistribution(uint32_t width, uint32_t height, bool isHdr, const void *texels, EnvDistribution &out);
} // namespace ptb
