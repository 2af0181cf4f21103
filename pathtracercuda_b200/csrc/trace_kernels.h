// trace_kernels.h — launch interface of the CUDA trace path (implemented in trace_kernels.cu).
#pragma once
#include "pt_types.h"
#include <cuda_runtime.h>

namespace ptb
{

// counters[] layout (device, unsigned long long)
enum { kCtrWork = 0, kCtrRays = 1, kCtrNodes = 2, kCtrPrims = 3, kCtrShades = 4, kCtrMisses = 5, kCtrError = 7,
       // scheduler statistics of the wavefront kernel (count_work = 1): stage executions and the slots they served
       kCtrTraceRounds = 8, kCtrTraceWalkers = 9, kCtrNodeIters = 10, kCtrShadeExec = 11, kCtrShadeSlots = 12, kCtrGenExec = 13, kCtrGenSlots = 14,
       kCtrLeafExec = 15, kCtrLeafSlots = 16, kCtrIdle = 17, kCtrBlocked = 18, kCtrRefills = 19, kCtrRefillSlots = 20, kCtrCount = 24 };

struct SceneDev
{
	const float4 *sceneBlob = nullptr; // [nodes | prims], 16-byte records, one allocation (bulk-copied to smem)
	const Mat *mats = nullptr;
	const TexDesc *textures = nullptr;
	uint32_t nodeCount = 0, primCount = 0, texCount = 0, skybox = 0;
	uint32_t globalCount = 0; // prims[0..globalCount): tested by every ray before the traversal (pt_types.h)
	uint32_t treeNodeCount = 0; // nodes[treeNodeCount..nodeCount): boxes of the hoisted primitives, for the pixel-beam walk only
};

struct RenderParams
{
	SceneDev scene;
	CameraDev cam;
	float4 *accum;
	unsigned long long *counters;
	uint32_t width, height, spp, ignoreHistory;
	uint32_t sampleOffset, sampleStride;
	uint32_t pixelOffset = 0, pixelStride = 1; // this launch renders pixels offset, offset + stride, ... (multi-GPU pixel partition)
	uint32_t seedLo, seedHi;
	uint32_t maxBounces;
	uint32_t regenLow = 1; // one-pixel-per-warp kernel: idle lanes wait until this many can start new samples together
	// one-pixel-per-warp kernel: the samples of a pixel are handed out in the order of their first scattering direction
	// (sortSamples, trace_kernels.cu).  sortScratch = sortStride x 4 bytes per warp of the grid; 0 = samples in index order
	uint16_t *sortScratch = nullptr;
	uint32_t sortStride = 0;
	uint32_t sortIgnore = 0; // timing aid: sort, then hand the samples out in index order all the same
	uint32_t sortBitsA = 4, sortBitsB = 2; // bins of the first / second random (2^(A+B) bins, 32..256)
	uint32_t beam = 0;     // one-pixel-per-warp kernel: camera rays take their leaves from the pixel's beam list (trace_device.cuh)
};

struct LaunchConfig
{
	int smCount = 148;
	int smemScene = 1;   // stage the scene in shared memory when it fits
	int countWork = 0;   // node/prim/shade/miss counters
	int variant = 0;     // kernel variant (0 = default: 12 for spp >= 64, else 4.  4: one pixel per lane, while-while traversal; 1: if/else traversal;
	                     //  5: + leaf parking; 8: one pixel per WARP (lanes = samples), while-while; 9/10: its other traversals;
	                     //  12: 8 with separate passes for camera rays and scattered rays;
	                     //  6: warp-pool wavefront, 7: CTA-pool warp-specialised wavefront - both measured slower)
	int traceLow = 0;    // warp-pool: run shade/generate early when fewer than this many lanes could traverse (0 = 24)
	int nodeLow = 0;     // warp-pool: the node loop leaves when fewer than this many lanes are still walking (0 = 24)
	int poolWarps = 0;   // warp-pool: warps per CTA (0 = as many as fit, <= 24)
	int traceWarps = 0;  // wavefront: warps per CTA that only traverse (0 = half of them); the others run the other stages
	int readyLow = -1;   // wavefront: stage warps run partial batches while the READY queue holds fewer rays than this (-1 = 128)
	int regenLow = 0;    // see RenderParams::regenLow (0 = default)
	int sortBitsA = 0, sortBitsB = -1; // 0 / -1 = defaults
	int sortSamples = -1; // order a pixel's samples by first scattering direction: 1 on, 0 off, -1 = on from 1024 spp (and spp <= 65535)
	int beam = -1;       // pixel beams for the camera rays of the one-pixel-per-warp kernel: 1 on, 0 off, -1 = on from 128 spp
	int poolSlots = 0;   // CTA-pool wavefront: path slots per CTA (0 = 1280 with the scene in shared memory, 1536 without)
	size_t maxSmemOptin = 0;
};

// Returns the number of kernels launched; *usedSmem = 1 when the scene was staged in shared memory.
int launchTrace(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem);
int launchTraceWarpPool(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem);   // trace_warppool.cu; 0 = not applicable
int launchTraceWavefront(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem); // trace_wavefront.cu; 0 = not applicable
int launchPrimary(const SceneDev &scene, const CameraDev &cam, uint32_t width, uint32_t height, int32_t *hitIndex, float *hitT, cudaStream_t stream);
int launchTraceRays(const SceneDev &scene, size_t n, const float *origins, const float *directions, float tMin, int32_t *hitIndex, float *hitT,
                    float *hitNormal, cudaStream_t stream);
int launchTonemap(const float4 *accum, uchar4 *out, uint32_t pixels, float invSampleCount, cudaStream_t stream);
int launchScale(const float4 *accum, float4 *out, uint32_t pixels, float scale, cudaStream_t stream);

} // namespace ptb
