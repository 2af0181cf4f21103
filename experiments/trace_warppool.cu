// trace_warppool.cu — the warp-pool WAVEFRONT kernel of the trace path (default).
//
// One persistent CTA per SM, 24 warps, no communication between warps.  Every warp owns a POOL of 64 path slots in
// shared memory (struct-of-arrays, 23 words per slot: ray, throughput, path radiance, pixel colour, pixel, sample/bounce
// state, leftover Philox words, closest hit, two deferred leaves) - twice as many paths as it has lanes - and four 64-bit
// masks in registers that say which pipeline stage each idle slot waits for:
//
//     gen    path ended (env miss, absorbed, max depth) or slot is new: env-miss lookup (trace.cu:115-134), accumulate
//            (:196-198), next sample / next pixel (warp-aggregated global atomic), camera ray (:187-192), test of the
//            scene-spanning primitives                                                            -> ready (or retired)
//     ready  has a ray: BVH traversal (hitBVH, trace.cu:28-98) in the lane's REGISTERS.  Leaves met on the way are
//            deferred (two per ray) so the loop contains nothing but the two-box node test       -> leaf | hit | gen
//     leaf   deferred primitive tests (Hittable::hit, Hittable.inl:88-145)                               -> hit | gen
//     hit    shade / sample (trace.cu:136-151) + scene-spanning primitives for the new ray          -> ready | gen
//
// Each iteration the warp runs ONE stage for up to 32 slots taken from that stage's mask (ballot/popc prefix compaction
// into a slot list), so generate / shade / intersect execute with well-filled warps whatever mixture of path depths
// and outcomes the pool holds.  The traversal stage leaves its node loop as soon as fewer than `nodeLow` lanes are still
// walking, retires the finished rays and refills those lanes from `ready`: it does not wait for the longest ray.
// A slot owns its pixel for all of its samples, added in order, so the image does not depend on scheduling
// (bit-identical to the per-lane kernels in trace_kernels.cu).
#include "trace_common.cuh"
#include <algorithm>

namespace ptb
{

constexpr int kPoolSlots = 64;
constexpr int kPoolThreads = 768;
enum : int { F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_TX, F_TY, F_TZ, F_LX, F_LY, F_LZ, F_CX, F_CY, F_CZ, F_PIXEL, F_STATE, F_RZ, F_RW, F_T, F_PRIM, F_PARK0, F_PARK1, kPoolWords };
constexpr uint32_t kStateSampleMask = 0x00ffffffu, kStateBounceShift = 24, kStateHasPath = 0x80000000u;
constexpr size_t kPoolBytesPerWarp = size_t(kPoolWords) * kPoolSlots * 4 + kPoolSlots; // + the slot list (one byte per entry)

struct PoolBest
{
	float t;
	int prim;
};
// closest-hit fold with the tie rule (equal t: larger scene index wins, Q7); the scene index of the current best is only
// fetched when a tie actually happens.  One out-of-line copy keeps the kernel inside the instruction cache.
template <bool SMEM>
__device__ __noinline__ PoolBest poolFoldPrim(const float4 *prims, uint32_t prim, V3 o, V3 d, PoolBest best)
{
	SceneView<SMEM> sv;
	sv.nodes = nullptr;
	sv.prims = prims;
	sv.globalCount = 0;
	const float4 *pp = prims + prim * 4;
	const float4 r0 = sv.ld(pp), r1 = sv.ld(pp + 1), r2 = sv.ld(pp + 2), meta = sv.ld(pp + 3);
	V3 lo, ld;
	toLocal(r0, r1, r2, o, d, lo, ld);
	float t;
	if (intersectLocal<kHotExact>(__float_as_uint(meta.x), lo, ld, 0.001f, best.t, t))
	{
		bool take = true;
		if (t == best.t && best.prim >= 0) take = !(__float_as_uint(meta.y) < __float_as_uint(sv.ld(prims + best.prim * 4 + 3).y));
		if (take) { best.t = t; best.prim = int(prim); }
	}
	return best;
}

template <bool SMEM, bool COUNT>
__global__ void __launch_bounds__(kPoolThreads, 1) traceKernelWP(const RenderParams p, const uint32_t sceneBytesAligned, const int traceLow, const int nodeLow)
{
	extern __shared__ __align__(128) float4 smemScene[];
	__shared__ uint64_t mbar;
	SceneView<SMEM> sv;
	if constexpr (SMEM)
	{
		stageSceneToSmem(smemScene, p.scene.sceneBlob, (p.scene.nodeCount + p.scene.primCount) * 64u, &mbar);
		sv.nodes = smemWindow(smemScene);
		sv.prims = sv.nodes + size_t(p.scene.nodeCount) * 4;
	}
	else
	{
		sv.nodes = p.scene.sceneBlob;
		sv.prims = p.scene.sceneBlob + size_t(p.scene.nodeCount) * 4;
	}
	sv.globalCount = p.scene.globalCount;
	constexpr uint32_t full = 0xffffffffu;
	constexpr int K = kPoolSlots;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;
	float *pool = reinterpret_cast<float *>(reinterpret_cast<char *>(smemScene) + sceneBytesAligned + size_t(warp) * kPoolBytesPerWarp);
	uint32_t *poolU = reinterpret_cast<uint32_t *>(pool);
	uint8_t *list = reinterpret_cast<uint8_t *>(pool + kPoolWords * K);
#define PF(field, slot) pool[(field) * K + (slot)]
#define PU(field, slot) poolU[(field) * K + (slot)]

	const uint32_t totalPixels = p.width * p.height;
	const float invW = 1.0f / float(p.width), invH = 1.0f / float(p.height);
	const V3 camO = mk(p.cam.origin[0], p.cam.origin[1], p.cam.origin[2]);
	constexpr float tMin = 0.001f;

	// all slots start in `gen` without a pixel
	PU(F_PIXEL, lane) = kInvalid; PU(F_PIXEL, lane + 32) = kInvalid;
	PU(F_STATE, lane) = p.spp & kStateSampleMask; PU(F_STATE, lane + 32) = p.spp & kStateSampleMask;
	PF(F_CX, lane) = 0.0f; PF(F_CY, lane) = 0.0f; PF(F_CZ, lane) = 0.0f;
	PF(F_CX, lane + 32) = 0.0f; PF(F_CY, lane + 32) = 0.0f; PF(F_CZ, lane + 32) = 0.0f;
	__syncwarp();
	unsigned long long ready = 0ull, hit = 0ull, gen = ~0ull, leaf = 0ull;

	// the lane's in-flight traversal
	int slot = -1;
	V3 ro = camO, rd = mk(0.0f, 0.0f, 1.0f);
	TravRay tr = makeTravRay(ro, rd);
	int cur = kEmptyChild, park0 = kEmptyChild, park1 = kEmptyChild, sp = 0, primBest = -1;
	float tBest = FLT_MAX;
	int stack[kStackSize];
	uint32_t rays = 0, nodeVisits = 0, primTests = 0, shades = 0, misses = 0;
	uint32_t dbg[13] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }; // lane 0: kCtrTraceRounds.. (count_work only)
#define DBG(idx, v) do { if (COUNT && lane == 0) dbg[(idx) - kCtrTraceRounds] += (v); } while (0)

	// slot ids of the set bits of m, ascending, into `list`; returns how many
	auto buildList = [&](unsigned long long m) -> int
	{
		const uint32_t lo = uint32_t(m), hi = uint32_t(m >> 32);
		const uint32_t cl = __popc(lo);
		if ((lo >> lane) & 1u) list[__popc(lo & ltMask)] = uint8_t(lane);
		if ((hi >> lane) & 1u) list[cl + __popc(hi & ltMask)] = uint8_t(lane + 32u);
		__syncwarp();
		return int(cl + __popc(hi));
	};
	// OR over the lanes of (pred ? bit s : 0)
	auto slotMask = [&](bool pred, int s) -> unsigned long long
	{
		const uint32_t lo = (pred && s < 32) ? (1u << s) : 0u, hi = (pred && s >= 32) ? (1u << (s - 32)) : 0u;
		return (unsigned long long)__reduce_or_sync(full, lo) | ((unsigned long long)__reduce_or_sync(full, hi) << 32);
	};
	auto foldLeaf = [&](int lf, V3 o, V3 d, float &tb, int &pb)
	{
		const uint32_t first = uint32_t(lf) & kLeafStartMask;
		const uint32_t count = (uint32_t(lf) >> kLeafCountShift) & 15u;
#pragma unroll 1
		for (uint32_t i = 0; i < count; ++i)
		{
			if (COUNT) ++primTests;
			PoolBest b;
			b.t = tb; b.prim = pb;
			b = poolFoldPrim<SMEM>(sv.prims, first + i, o, d, b);
			tb = b.t; pb = b.prim;
		}
	};
	// the scene-spanning primitives, tested where the warp is full and every lane reads the same record
	auto foldGlobals = [&](V3 o, V3 d, int s)
	{
		PoolBest gb;
		gb.t = FLT_MAX; gb.prim = -1;
#pragma unroll 1
		for (uint32_t g = 0; g < sv.globalCount; ++g)
		{
			if (COUNT) ++primTests;
			gb = poolFoldPrim<SMEM>(sv.prims, g, o, d, gb);
		}
		PF(F_T, s) = gb.t; PU(F_PRIM, s) = uint32_t(gb.prim);
	};
	// defer the leaves the walk arrives at and keep walking (the sentinel is never deferred: it ends the walk)
	auto parkLeaves = [&]()
	{
		while (cur < 0 && cur != kEmptyChild)
		{
			if (park0 == kEmptyChild) park0 = cur;
			else if (park1 == kEmptyChild) park1 = cur;
			else break; // a third leaf: the lane stops walking until the warp tests what it holds
			cur = stack[--sp];
		}
	};

	while (true)
	{
		const int nIn = __popc(__ballot_sync(full, slot >= 0));
		const int nReady = __popcll(ready), nHit = __popcll(hit), nGen = __popcll(gen), nLeaf = __popcll(leaf);
		if (nIn + nReady + nHit + nGen + nLeaf == 0) break;
		const int traceAvail = min(nIn + nReady, 32);
		// 0 trace, 1 shade, 2 generate, 3 deferred leaves.  Full stages first; traverse while enough lanes can be kept
		// busy; otherwise the fullest of the other stages
		int stage = 0;
		if (nHit >= 32) stage = 1;
		else if (nGen >= 32) stage = 2;
		else if (nLeaf >= 32) stage = 3;
		else if (traceAvail < traceLow && (nHit | nGen | nLeaf) != 0)
		{
			if (nLeaf >= nHit && nLeaf >= nGen) stage = 3;
			else if (nHit >= nGen) stage = 1;
			else stage = 2;
		}

		if (stage == 3)
		{
			// ---------------- deferred primitive tests (Hittable::hit, Hittable.inl:88-145) ----------------
			const int n = min(buildList(leaf), 32);
			DBG(kCtrLeafExec, 1); DBG(kCtrLeafSlots, n);
			const bool mine = int(lane) < n;
			const int s = mine ? int(list[lane]) : 0;
			const unsigned long long taken = n == nLeaf ? leaf : slotMask(mine, s);
			leaf &= ~taken;
			int lprim = -1;
			if (mine)
			{
				const V3 lro = mk(PF(F_OX, s), PF(F_OY, s), PF(F_OZ, s)), lrd = mk(PF(F_DX, s), PF(F_DY, s), PF(F_DZ, s));
				float lt = PF(F_T, s);
				lprim = int(PU(F_PRIM, s));
				foldLeaf(int(PU(F_PARK0, s)), lro, lrd, lt, lprim);
				foldLeaf(int(PU(F_PARK1, s)), lro, lrd, lt, lprim); // kEmptyChild has count 0
				PF(F_T, s) = lt; PU(F_PRIM, s) = uint32_t(lprim);
			}
			const unsigned long long hitMask = slotMask(mine && lprim >= 0, s);
			hit |= hitMask;
			gen |= taken & ~hitMask;
			__syncwarp();
			continue;
		}

		if (stage == 1)
		{
			// ---------------- shade / sample stage (trace.cu:136-151) ----------------
			const int n = min(buildList(hit), 32);
			DBG(kCtrShadeExec, 1); DBG(kCtrShadeSlots, n);
			const bool mine = int(lane) < n;
			const int s = mine ? int(list[lane]) : 0;
			const unsigned long long taken = n == nHit ? hit : slotMask(mine, s);
			hit &= ~taken;
			bool cont = false;
			if (mine)
			{
				if (COUNT) ++shades;
				const V3 sro = mk(PF(F_OX, s), PF(F_OY, s), PF(F_OZ, s)), srd = mk(PF(F_DX, s), PF(F_DY, s), PF(F_DZ, s));
				const int prim = int(PU(F_PRIM, s));
				const Surface sf = surfaceAt<SMEM>(sv, prim, sro, srd, PF(F_T, s));
				const float4 *mp = reinterpret_cast<const float4 *>(p.scene.mats + prim);
				const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
				V3 thr = mk(PF(F_TX, s), PF(F_TY, s), PF(F_TZ, s));
				const V3 L = mk(PF(F_LX, s), PF(F_LY, s), PF(F_LZ, s)) + thr * mk(m1.x, m1.y, m1.z); // getEmitted, Material.inl:62-65
				PF(F_LX, s) = L.x; PF(F_LY, s) = L.y; PF(F_LZ, s) = L.z;
				V3 base = mk(m0.x, m0.y, m0.z);
				const uint32_t tex = __float_as_uint(m2.x), mtype = __float_as_uint(m2.y);
				if (tex != 0 && tex <= p.scene.texCount)
				{
					const V3 tap = texLookupNI(p.scene.textures, tex, sf.u, sf.v); // Material.inl:26-35
					base = mk(fastPow(tap.x, 2.2f), fastPow(tap.y, 2.2f), fastPow(tap.z, 2.2f));
				}
				const uint32_t state = PU(F_STATE, s);
				uint32_t bounce = (state >> kStateBounceShift) & 0x7fu;
				float rnd0, rnd1;
				if (bounce != 0u && (bounce & 1u))
				{
					const uint4 r = philoxNI(PU(F_PIXEL, s), p.sampleOffset + (state & kStateSampleMask) * p.sampleStride, (bounce + 1u) >> 1, p.seedLo, p.seedHi);
					rnd0 = uniform01(r.x); rnd1 = uniform01(r.y);
					PU(F_RZ, s) = r.z; PU(F_RW, s) = r.w;
				}
				else { rnd0 = uniform01(PU(F_RZ, s)); rnd1 = uniform01(PU(F_RW, s)); }
				V3 wi, weight;
				cont = sampleMaterial(mtype, base, m0.w, m1.w, sf.n, srd, rnd0, rnd1, wi, weight);
				if (cont)
				{
					++bounce;
					if (bounce >= p.maxBounces) cont = false;
				}
				if (cont)
				{
					thr = thr * weight;
					PF(F_TX, s) = thr.x; PF(F_TY, s) = thr.y; PF(F_TZ, s) = thr.z;
					PF(F_OX, s) = sf.p.x; PF(F_OY, s) = sf.p.y; PF(F_OZ, s) = sf.p.z;
					PF(F_DX, s) = wi.x; PF(F_DY, s) = wi.y; PF(F_DZ, s) = wi.z;
					PU(F_STATE, s) = (state & ~(0x7fu << kStateBounceShift)) | (bounce << kStateBounceShift);
					foldGlobals(sf.p, wi, s);
				}
			}
			const unsigned long long contMask = slotMask(cont, s);
			ready |= contMask;
			gen |= taken & ~contMask; // F_PRIM stays >= 0: the generate stage will not look up the environment
			__syncwarp();
			continue;
		}

		if (stage == 2)
		{
			// ---------------- env-miss + accumulate + generate stage (trace.cu:115-134, :187-198) ----------------
			const int n = min(buildList(gen), 32);
			DBG(kCtrGenExec, 1); DBG(kCtrGenSlots, n);
			const bool mine = int(lane) < n;
			const int s = mine ? int(list[lane]) : 0;
			const unsigned long long taken = n == nGen ? gen : slotMask(mine, s);
			gen &= ~taken;
			uint32_t pixel = kInvalid, sample = 0;
			V3 color = mk(0.0f, 0.0f, 0.0f);
			bool need = false;
			if (mine)
			{
				const uint32_t state = PU(F_STATE, s);
				pixel = PU(F_PIXEL, s);
				sample = state & kStateSampleMask;
				color = mk(PF(F_CX, s), PF(F_CY, s), PF(F_CZ, s));
				if (state & kStateHasPath)
				{
					V3 L = mk(PF(F_LX, s), PF(F_LY, s), PF(F_LZ, s));
					if (int(PU(F_PRIM, s)) < 0)
					{
						if (COUNT) ++misses;
						if (p.scene.skybox != 0)
						{
							const V3 mrd = mk(PF(F_DX, s), PF(F_DY, s), PF(F_DZ, s)), thr = mk(PF(F_TX, s), PF(F_TY, s), PF(F_TZ, s));
							const float theta = fastAcos(mrd.y), phi = fastAtan2(mrd.z, mrd.x);
							const V3 sky = texLookupNI(p.scene.textures, p.scene.skybox, phi * (0.5f / PT_PI), theta * (1.0f / PT_PI));
							L = L + thr * sky;
						}
					}
					color = color + L;
					++sample;
				}
				need = sample >= p.spp;
				if (need && pixel != kInvalid)
				{
					float4 out = make_float4(color.x, color.y, color.z, 1.0f); // trace.cu:196-198
					if (!p.ignoreHistory)
					{
						const float4 prev = p.accum[pixel];
						out.x += prev.x; out.y += prev.y; out.z += prev.z;
					}
					p.accum[pixel] = out;
				}
			}
			bool alive = mine;
			const uint32_t needMask = __ballot_sync(full, need);
			if (needMask)
			{
				const uint32_t leader = __ffs(needMask) - 1;
				unsigned long long base = 0;
				if (lane == leader) base = atomicAdd(&p.counters[kCtrWork], (unsigned long long)__popc(needMask));
				base = __shfl_sync(full, base, leader);
				if (need)
				{
					const unsigned long long m = base + __popc(needMask & ltMask);
					if (m >= totalPixels) alive = false; // slot retires: in no mask from now on
					else { pixel = uint32_t(m); sample = 0; color = mk(0.0f, 0.0f, 0.0f); }
				}
			}
			if (alive)
			{
				const uint32_t sampleIdx = p.sampleOffset + sample * p.sampleStride;
				const uint4 r = philoxNI(pixel, sampleIdx, 0u, p.seedLo, p.seedHi);
				uint32_t px, py;
				pixelToXY(pixel, p.width, p.height, px, py);
				const float u = (float(px) + uniform01(r.x)) * invW; // trace.cu:190
				const float v = (float(py) + uniform01(r.y)) * invH;
				const V3 d = cameraDir<kHotExact>(p.cam, u, v);
				PF(F_OX, s) = camO.x; PF(F_OY, s) = camO.y; PF(F_OZ, s) = camO.z;
				PF(F_DX, s) = d.x; PF(F_DY, s) = d.y; PF(F_DZ, s) = d.z;
				PF(F_TX, s) = 1.0f; PF(F_TY, s) = 1.0f; PF(F_TZ, s) = 1.0f;
				PF(F_LX, s) = 0.0f; PF(F_LY, s) = 0.0f; PF(F_LZ, s) = 0.0f;
				PF(F_CX, s) = color.x; PF(F_CY, s) = color.y; PF(F_CZ, s) = color.z;
				PU(F_RZ, s) = r.z; PU(F_RW, s) = r.w;
				PU(F_PIXEL, s) = pixel;
				PU(F_STATE, s) = kStateHasPath | sample;
				foldGlobals(camO, d, s);
			}
			ready |= slotMask(alive, s);
			__syncwarp();
			continue;
		}

		// ---------------- BVH-traverse stage (trace.cu:112, hitBVH :28-98) ----------------
		{
			const uint32_t needRay = __ballot_sync(full, slot < 0);
			if (needRay != 0u && ready != 0ull)
			{
				const int nList = buildList(ready);
				const int r = __popc(needRay & ltMask);
				const bool take = slot < 0 && r < nList;
				DBG(kCtrRefills, 1); DBG(kCtrRefillSlots, min(nList, __popc(needRay)));
				if (take)
				{
					slot = int(list[r]);
					ro = mk(PF(F_OX, slot), PF(F_OY, slot), PF(F_OZ, slot));
					rd = mk(PF(F_DX, slot), PF(F_DY, slot), PF(F_DZ, slot));
					tr = makeTravRay(ro, rd);
					tBest = PF(F_T, slot);
					primBest = int(PU(F_PRIM, slot));
					stack[0] = kEmptyChild; // sentinel: a leaf reference with zero primitives
					sp = 1; cur = 0; park0 = kEmptyChild; park1 = kEmptyChild;
					++rays;
				}
				ready &= ~slotMask(take, slot);
				__syncwarp();
			}
		}
		// node loop: leaves when fewer than `target` lanes are still walking
		const uint32_t walkersStart = __popc(__ballot_sync(full, cur >= 0));
		const uint32_t target = max(min(uint32_t(nodeLow), walkersStart), 1u);
		DBG(kCtrTraceRounds, 1); DBG(kCtrTraceWalkers, walkersStart);
		while (__popc(__ballot_sync(full, cur >= 0)) >= target)
		{
			DBG(kCtrNodeIters, 1);
			if (cur >= 0)
			{
				if (COUNT) ++nodeVisits;
				const float4 *n = sv.nodes + cur * 4;
				const float4 A = sv.ld(n), Bq = sv.ld(n + 1), C = sv.ld(n + 2);
				const float4 Dq = sv.ld(n + 3);
				bool hitA, hitB;
				float nearA, nearB;
				testNodeBoxes(A, Bq, C, tr, tMin, tBest, hitA, hitB, nearA, nearB);
				const int cA = __float_as_int(Dq.x), cB = __float_as_int(Dq.y);
				if (hitA && hitB)
				{
					const bool bFirst = nearB < nearA;
					stack[sp++] = bFirst ? cA : cB;
					cur = bFirst ? cB : cA;
				}
				else if (hitA) cur = cA;
				else if (hitB) cur = cB;
				else cur = stack[--sp];
				parkLeaves();
			}
		}
		// lanes holding a third leaf: test everything they hold now (rare in sparse scenes, the rule in dense ones)
		const bool blocked = slot >= 0 && cur < 0 && cur != kEmptyChild;
		if (__any_sync(full, blocked))
		{
			DBG(kCtrBlocked, 1);
			if (blocked)
			{
				foldLeaf(park0, ro, rd, tBest, primBest);
				foldLeaf(park1, ro, rd, tBest, primBest);
				foldLeaf(cur, ro, rd, tBest, primBest);
				park0 = kEmptyChild; park1 = kEmptyChild;
				cur = stack[--sp];
				parkLeaves();
			}
		}
		// retire finished traversals
		const bool done = slot >= 0 && cur == kEmptyChild;
		if (__any_sync(full, done))
		{
			const bool needLeaf = done && park0 != kEmptyChild;
			if (done)
			{
				PF(F_T, slot) = tBest; PU(F_PRIM, slot) = uint32_t(primBest);
				PU(F_PARK0, slot) = uint32_t(park0); PU(F_PARK1, slot) = uint32_t(park1);
			}
			const unsigned long long fin = slotMask(done, slot), finLeaf = slotMask(needLeaf, slot), finHit = slotMask(done && !needLeaf && primBest >= 0, slot);
			leaf |= finLeaf;
			hit |= finHit;
			gen |= fin & ~(finLeaf | finHit);
			if (done) slot = -1;
			__syncwarp();
		}
	}
#undef PF
#undef PU

	unsigned long long r64 = rays;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) r64 += __shfl_xor_sync(full, r64, o);
	if (lane == 0) atomicAdd(&p.counters[kCtrRays], r64);
	if (COUNT)
	{
		unsigned long long c[4] = { nodeVisits, primTests, shades, misses };
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(full, c[k], o);
			if (lane == 0) atomicAdd(&p.counters[kCtrNodes + k], c[k]);
		}
		if (lane == 0)
			for (int k = 0; k < 13; ++k) atomicAdd(&p.counters[kCtrTraceRounds + k], (unsigned long long)dbg[k]);
	}
#undef DBG
}

template <typename K>
static int launchPool(K kern, const RenderParams &p, const LaunchConfig &cfg, size_t sceneBytesAligned, int warps, cudaStream_t stream)
{
	const size_t smemBytes = sceneBytesAligned + size_t(warps) * kPoolBytesPerWarp;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemBytes));
	const int traceLow = cfg.traceLow > 0 ? cfg.traceLow : 24;
	const int nodeLow = cfg.nodeLow > 0 ? cfg.nodeLow : 24;
	kern<<<cfg.smCount, warps * 32, smemBytes, stream>>>(p, uint32_t(sceneBytesAligned), traceLow, nodeLow);
	return 1;
}

// Returns 0 when the kernel cannot run this configuration (the caller falls back to the per-lane kernel).
int launchTraceWarpPool(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem)
{
	if (p.spp > kStateSampleMask || p.maxBounces > 0x7fu) return 0;
	// one CTA per SM, as many warps as the pools (and the scene copy) leave room for
	const size_t sceneBytes = (size_t(p.scene.nodeCount) + p.scene.primCount) * 64;
	const size_t aligned = (sceneBytes + 127) & ~size_t(127);
	const size_t avail = cfg.maxSmemOptin > 2048 ? cfg.maxSmemOptin - 2048 : 0;
	const int maxWarps = cfg.poolWarps > 0 ? std::min(cfg.poolWarps, kPoolThreads / 32) : kPoolThreads / 32;
	const bool smem = cfg.smemScene && aligned + size_t(16) * kPoolBytesPerWarp <= avail;
	const int warps = int(std::min<size_t>(size_t(maxWarps), (avail - (smem ? aligned : 0)) / kPoolBytesPerWarp));
	if (warps < 1) return 0;
	if (usedSmem) *usedSmem = smem ? 1 : 0;
	if (smem) return cfg.countWork ? launchPool(traceKernelWP<true, true>, p, cfg, aligned, warps, stream) : launchPool(traceKernelWP<true, false>, p, cfg, aligned, warps, stream);
	return cfg.countWork ? launchPool(traceKernelWP<false, true>, p, cfg, 0, warps, stream) : launchPool(traceKernelWP<false, false>, p, cfg, 0, warps, stream);
}

} // namespace ptb
