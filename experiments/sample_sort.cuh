// The round-1 sample sort of the one-pixel-per-warp kernel (superseded by first-bounce stratification; not built).
// It was called once per pixel from traceKernel and its order read back through RenderParams::sortScratch when samples were handed out.
// ---------------------------------------------------------------------------------------------------------------
// Sample order.  The lanes of a warp trace samples of one pixel; the scattered rays of the first bounce leave (nearly) the
// same point, in directions set by the sample's first two BSDF randoms.  Handing the samples out in the order of those
// randoms - 32..256 bins: lobe + azimuth (the top bits of the first), polar angle (the top bits of the second), neighbours adjacent -
// puts similar directions into the same pass: similar walks, the same leaves, the same hit-or-miss outcome, the same lobe.
// It is only an order: every sample is still drawn from its own Philox counter, so the estimator and the set of paths are
// unchanged (the image differs by float summation order alone).  A counting sort by the warp, once per pixel: pass A
// draws each sample's randoms and counts the bins (shared-memory atomics: only the counts are used), pass B gives every
// sample its place (ranks from match_any, not from atomics, so the order is the same from run to run).  `order` and `keys` live in global scratch (L2).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSortBinsMax = 128; // bins = 2^(bitsA + bitsB) <= 128, >= 32
#ifdef PTB_NO_SORT
constexpr bool kSortCompiled = false;
#else
constexpr bool kSortCompiled = true;
#endif
// Scratch of one warp (global memory, L2): order[stride] uint16 | keys[stride] uint16.  Four samples per lane and iteration
// (s = base + 4 * lane + k): the four Philox evaluations are independent, the key loads and stores are vectors.
// (Keeping the draws too, so that generating the camera ray reads slot 0 back instead of drawing it again, was measured:
// 634 vs 574 ms - a dependent 16-byte gather from a 300 MB scratch at the head of every camera pass.)
static __device__ __forceinline__ size_t sortScratchBytesPerWarp(uint32_t stride) { return size_t(stride) * (2u + 2u); }
static __device__ __forceinline__ char *sortScratchOfWarp(const RenderParams &p)
{
	return reinterpret_cast<char *>(p.sortScratch) + (size_t(blockIdx.x) * (kTraceThreads / 32) + (threadIdx.x >> 5)) * sortScratchBytesPerWarp(p.sortStride);
}
// (two arguments on purpose: the call sits in the kernel's main loop, and every argument register is one the loop loses)
static __device__ __noinline__ void sortSamples(uint32_t pixel, const RenderParams *pp, uint32_t *hist)
{
	const RenderParams &p = *pp;
	const uint32_t spp = p.spp, sampleOffset = p.sampleOffset, sampleStride = p.sampleStride, seedLo = p.seedLo, seedHi = p.seedHi;
	const uint32_t bitsA = p.sortBitsA, bitsB = p.sortBitsB;
	char *mine = sortScratchOfWarp(p);
	uint16_t *order = reinterpret_cast<uint16_t *>(mine);
	uint16_t *keys = order + p.sortStride;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t nbA = bitsA & 15u;
	const uint32_t bins = 1u << (nbA + bitsB), perLane = bins >> 5;
	for (uint32_t k = lane; k < bins; k += 32) hist[k] = 0;
	__syncwarp();
	// pass A: draws, keys, histogram
	for (uint32_t base = 0; base < spp; base += 128)
	{
		const uint32_t s0 = base + 4u * lane;
		uint32_t key[4];
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			const uint32_t s = s0 + uint32_t(k);
			key[k] = bins; // samples past the end: a bin of their own, never counted
			if (s < spp)
			{
				const uint4 r = philox4x32_10(pixel, sampleOffset + s * sampleStride, 0u, 0u, seedLo, seedHi);
				const uint32_t a = r.z >> (32u - nbA), bb = bitsB ? r.w >> (32u - bitsB) : 0u;
				// snake through the minor bins: neighbouring keys are neighbouring directions
				if (bitsA & 16u) key[k] = (bb << nbA) | ((bb & 1u) ? ((1u << nbA) - 1u) - a : a);
				else key[k] = (a << bitsB) | ((a & 1u) ? ((1u << bitsB) - 1u) - bb : bb);
			}
		}
		if (s0 < spp) __stcg(reinterpret_cast<uint2 *>(keys + s0), make_uint2(key[0] | (key[1] << 16), key[2] | (key[3] << 16))); // stride is a multiple of 128
#pragma unroll
		for (int k = 0; k < 4; ++k)
			if (key[k] < bins) atomicAdd(&hist[key[k]], 1u); // only the COUNTS are used: no order dependence
	}
	__syncwarp();
	// counts -> first position of every bin (lane l owns bins l * perLane ... + perLane - 1)
	{
		uint32_t sum = 0;
		for (uint32_t k = 0; k < perLane; ++k) sum += hist[lane * perLane + k];
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (int(lane) >= o) incl += t;
		}
		uint32_t run = incl - sum;
		__syncwarp();
		for (uint32_t k = 0; k < perLane; ++k)
		{
			const uint32_t c = hist[lane * perLane + k];
			hist[lane * perLane + k] = run;
			run += c;
		}
		__syncwarp();
	}
	// pass B: places.  Stable in the order (k, lane) within a 128-sample batch - any fixed order will do, it is the same from run to run
	for (uint32_t base = 0; base < spp; base += 128)
	{
		const uint32_t s0 = base + 4u * lane;
		uint2 packed = make_uint2(bins | (bins << 16), bins | (bins << 16));
		if (s0 < spp) packed = __ldcg(reinterpret_cast<const uint2 *>(keys + s0));
		const uint32_t key[4] = { packed.x & 0xffffu, packed.x >> 16, packed.y & 0xffffu, packed.y >> 16 };
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			const uint32_t same = __match_any_sync(0xffffffffu, key[k]);
			uint32_t first = 0;
			if (key[k] < bins) first = hist[key[k]];
			__syncwarp();
			if (key[k] < bins)
			{
				order[first + uint32_t(__popc(same & lt))] = uint16_t(s0 + uint32_t(k));
				if ((same & lt) == 0u) hist[key[k]] = first + uint32_t(__popc(same));
			}
			__syncwarp();
		}
	}
	__threadfence_block();
	__syncwarp();
}

