// trace_common.cuh — pieces shared by the trace kernels (trace_kernels.cu).
#pragma once
#define PTB_PRIM_FN __device__ __noinline__ // one out-of-line primitive test per kernel (code size, see trace_device.cuh)
#define PTB_BEAM_FN __device__ __noinline__ // once per pixel: kept out of the hot loop's code
#define PTB_ATAN_FN static __device__ __noinline__ // the texture-coordinate call sites (sphere / cylinder UV) share one copy; the sky lookup of the hot loop inlines its own
#include "trace_device.cuh"

namespace ptb
{

constexpr int kThreads = 256;
// the trace kernel runs one CTA of 32 warps per SM: 64 registers per thread (no spills), ONE shared-memory copy of the
// scene per SM, the rest of the 228 KB stays L1 for the traversal stack / materials / textures (measured on
// generated_scene: 3 x 256 threads 12.3, 2 x 512 12.9, 1 x 1024 13.2 Grays/s)
#ifndef PTB_TRACE_THREADS
#define PTB_TRACE_THREADS 1024
#endif
constexpr int kTraceThreads = PTB_TRACE_THREADS;
static_assert(kTraceThreads % 32 == 0 && kTraceThreads * 4 <= int(kStackStride), "the shared-memory stack holds one int per thread and level");

// pixel index -> (x, y) without an integer division (exact while the index is exact in fp32; one correction step)
__device__ __forceinline__ void pixelToXY(uint32_t pixel, uint32_t width, uint32_t height, uint32_t &px, uint32_t &py)
{
	if (width * height <= (1u << 24))
	{
		py = __float2uint_rz(__fdividef(__uint2float_rn(pixel) + 0.5f, __uint2float_rn(width)));
		px = pixel - py * width;
		if (int(px) < 0) { --py; px += width; }
		else if (px >= width) { ++py; px -= width; }
	}
	else { px = pixel % width; py = pixel / width; }
}

constexpr uint32_t kInvalid = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------------
// TMA bulk copy of the scene blob into shared memory (cp.async.bulk + mbarrier), one elected thread per CTA
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stageSceneToSmem(void *smemDst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
	const uint32_t mbarAddr = uint32_t(__cvta_generic_to_shared(mbar));
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbarAddr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbarAddr), "r"(bytes) : "memory");
		uint32_t done = 0;
		const uint32_t dst = uint32_t(__cvta_generic_to_shared(smemDst));
		while (done < bytes)
		{
			const uint32_t chunk = min(bytes - done, 32768u);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + done),
			             "l"(reinterpret_cast<const char *>(gsrc) + done), "r"(chunk), "r"(mbarAddr)
			             : "memory");
			done += chunk;
		}
	}
	// everyone waits for phase 0 of the barrier to complete (all bytes landed)
	uint32_t ready = 0;
	while (!ready)
	{
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ready) : "r"(mbarAddr), "r"(0u) : "memory");
	}
}


} // namespace ptb
