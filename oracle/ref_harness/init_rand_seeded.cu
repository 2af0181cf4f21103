// Second-seed RNG initialisation for the noise-floor render.
// The reference hard-codes its XORWOW seed base (kernels/initRandState.cu:16, "1984 + pixel").  To get an
// INDEPENDENT-seed reference render without patching reference sources, oracle/_ref/ref_pt_seedB links every
// reference translation unit EXCEPT kernels/initRandState.cu and takes this definition of the same symbol.
// Harness file, ours.
#include "pathtracer/kernels/initRandState.h"
#include <cstdint>
#ifndef SEED_BASE
#define SEED_BASE 7919
#endif
__global__ void initRandState(int width, int height, curandState *randState)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x < width && y < height)
	{
		const uint32_t pixel = uint32_t(y) * uint32_t(width) + uint32_t(x);
		curand_init(SEED_BASE + pixel, 0, 0, &randState[pixel]);
	}
}
