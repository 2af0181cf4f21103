// trace_kernels.cu — the trace hot path as hand-written sm_100a CUDA.
//
// Replaces the reference's three kernels (kernels/trace.cu traceKernel :158-199 + getColor :101-156 + hitBVH :28-98,
// kernels/initRandState.cu, kernels/tonemap.cu) and the device halves of Hittable.inl / Material.inl / MonteCarlo.h /
// brdf.h / Camera.inl.  Same estimator, different machine mapping:
//
//   * persistent CTAs (one wave sized to the SM count); every lane owns a pixel and regenerates a new path the moment
//     its current one terminates, so warps stay full through all five path segments instead of idling until the
//     longest path of the warp ends (the reference runs one thread per pixel with no compaction);
//   * pixels are handed out by a warp-aggregated atomic (ballot + popc prefix) - no per-pixel RNG state, no 48 B/pixel
//     buffer: Philox4x32-10 keyed on (seed), counter (pixel, sample, bounce slot);
//   * the BVH is 64-byte two-box nodes + 64-byte primitives read with 128-bit loads; when nodes+primitives fit in
//     shared memory they are staged there once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier);
//   * ray/box tests use a precomputed reciprocal direction (12 FMA per node instead of the reference's 3 IEEE
//     divisions per box, AABB.inl:26); the four quadrics share one intersection routine selected by coefficients, the
//     two flat shapes share another: 3 divergent classes instead of 7;
//   * normals, UVs, the tangent frame and the material fetch happen once per path segment at the closest hit, not once
//     per accepted candidate (Hittable.inl:129-142).
#include "trace_device.cuh"

namespace ptb
{

constexpr int kThreads = 256;
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

// ---------------------------------------------------------------------------------------------------------------
// the trace kernel
// ---------------------------------------------------------------------------------------------------------------
template <bool SMEM, bool COUNT, int TRAV>
__global__ void __launch_bounds__(kThreads, 3) traceKernel(const RenderParams p)
{
	extern __shared__ __align__(128) float4 smemScene[];
	__shared__ uint64_t mbar;
	SceneView<SMEM> sv;
	if constexpr (SMEM)
	{
		stageSceneToSmem(smemScene, p.scene.sceneBlob, (p.scene.nodeCount + p.scene.primCount) * 64u, &mbar);
		sv.nodes = smemScene;
		sv.prims = smemScene + size_t(p.scene.nodeCount) * 4;
	}
	else
	{
		sv.nodes = p.scene.sceneBlob;
		sv.prims = p.scene.sceneBlob + size_t(p.scene.nodeCount) * 4;
	}

	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t totalPixels = p.width * p.height;
	const V3 camO = mk(p.cam.origin[0], p.cam.origin[1], p.cam.origin[2]);

	uint32_t pixel = kInvalid, sample = p.spp;
	bool active = true, alive = false;
	V3 color = mk(0.0f, 0.0f, 0.0f);
	V3 ro = camO, rd = mk(0.0f, 0.0f, 1.0f), thr = mk(1.0f, 1.0f, 1.0f), L = mk(0.0f, 0.0f, 0.0f);
	uint32_t bounce = 0, rz = 0, rw = 0, sampleIdx = 0;
	uint32_t rays = 0, nodeVisits = 0, primTests = 0, shades = 0, misses = 0;

	while (true)
	{
		// ---- accumulate + fetch the next pixel (warp-aggregated atomic: ballot + popc prefix) ----
		const bool need = active && !alive && sample == p.spp;
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
		const uint32_t needMask = __ballot_sync(0xffffffffu, need);
		if (needMask)
		{
			const uint32_t leader = __ffs(needMask) - 1;
			unsigned long long base = 0;
			if (lane == leader) base = atomicAdd(&p.counters[kCtrWork], (unsigned long long)__popc(needMask));
			base = __shfl_sync(0xffffffffu, base, leader);
			if (need)
			{
				const unsigned long long mine = base + __popc(needMask & ((1u << lane) - 1u));
				if (mine >= totalPixels) { active = false; pixel = kInvalid; }
				else { pixel = uint32_t(mine); sample = 0; color = mk(0.0f, 0.0f, 0.0f); }
			}
		}
		if (!__any_sync(0xffffffffu, active)) break;

		if (active)
		{
			// ---- generate (trace.cu:187-192) ----
			if (!alive)
			{
				sampleIdx = p.sampleOffset + sample * p.sampleStride;
				const uint4 r = philox4x32_10(pixel, sampleIdx, 0u, 0u, p.seedLo, p.seedHi);
				const uint32_t px = pixel % p.width, py = pixel / p.width;
				const float u = divExact(float(px) + uniform01(r.x), float(p.width)); // trace.cu:190
				const float v = divExact(float(py) + uniform01(r.y), float(p.height));
				rz = r.z; rw = r.w;
				ro = camO;
				rd = cameraDir(p.cam, u, v);
				thr = mk(1.0f, 1.0f, 1.0f);
				L = mk(0.0f, 0.0f, 0.0f);
				bounce = 0;
				alive = true;
			}

			// ---- traverse + intersect (trace.cu:112) ----
			++rays;
			const Hit h = TRAV == 0 ? closestHit<SMEM, COUNT>(sv, ro, rd, 0.001f, nodeVisits, primTests)
			              : TRAV == 1 ? closestHitWW<SMEM, COUNT, false>(sv, ro, rd, 0.001f, nodeVisits, primTests)
			                          : closestHitWW<SMEM, COUNT, true>(sv, ro, rd, 0.001f, nodeVisits, primTests);

			bool terminate;
			if (h.prim < 0)
			{
				// ---- environment miss (trace.cu:115-134) ----
				if (COUNT) ++misses;
				if (p.scene.skybox != 0)
				{
					const float theta = acosf(rd.y), phi = atan2f(rd.z, rd.x);
					const V3 sky = texLookup(p.scene.textures, p.scene.skybox, phi / (2.0f * PT_PI), theta / PT_PI);
					L = L + thr * sky;
				}
				terminate = true;
			}
			else
			{
				// ---- shade / sample (trace.cu:136-151) ----
				if (COUNT) ++shades;
				const Surface s = surfaceAt<SMEM>(sv, h.prim, ro, rd, h.t);
				const float4 *mp = reinterpret_cast<const float4 *>(p.scene.mats + h.prim);
				const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
				L = L + thr * mk(m1.x, m1.y, m1.z); // getEmitted, Material.inl:62-65
				V3 base = mk(m0.x, m0.y, m0.z);
				const uint32_t tex = __float_as_uint(m2.x), mtype = __float_as_uint(m2.y);
				if (tex != 0 && tex <= p.scene.texCount)
				{
					const V3 tap = texLookup(p.scene.textures, tex, s.u, s.v); // Material.inl:26-35
					base = mk(fastPow(tap.x, 2.2f), fastPow(tap.y, 2.2f), fastPow(tap.z, 2.2f));
				}
				float rnd0, rnd1;
				if (bounce == 0) { rnd0 = uniform01(rz); rnd1 = uniform01(rw); }
				else if (bounce & 1u)
				{
					const uint4 r = philox4x32_10(pixel, sampleIdx, (bounce + 1u) >> 1, 0u, p.seedLo, p.seedHi);
					rnd0 = uniform01(r.x); rnd1 = uniform01(r.y);
					rz = r.z; rw = r.w;
				}
				else { rnd0 = uniform01(rz); rnd1 = uniform01(rw); }
				V3 wi, weight;
				const bool cont = sampleMaterial(mtype, base, m0.w, m1.w, s.n, rd, rnd0, rnd1, wi, weight);
				terminate = !cont;
				if (cont)
				{
					thr = thr * weight;
					ro = s.p;
					rd = wi;
					++bounce;
					if (bounce >= p.maxBounces) terminate = true;
				}
			}
			if (terminate)
			{
				color = color + L;
				++sample;
				alive = false;
			}
		}
	}

	// ---- counters: one atomic per warp ----
	unsigned long long r64 = rays;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) r64 += __shfl_xor_sync(0xffffffffu, r64, o);
	if (lane == 0) atomicAdd(&p.counters[kCtrRays], r64);
	if (COUNT)
	{
		unsigned long long c[4] = { nodeVisits, primTests, shades, misses };
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
			if (lane == 0) atomicAdd(&p.counters[kCtrNodes + k], c[k]);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// traceKernelV2 — warp-level wavefront.  Same per-path arithmetic as traceKernel (so the image is bit-identical), but
// the warp no longer lets every lane run its own control flow.  Each lane is a small state machine whose next step is
// one of six STAGES:
//     NODE   one two-box BVH node test            PQ / PF / PC   one primitive test of the quadric / flat / cube class
//     SHADE  surface + material sample at a hit    GEN            env-miss lookup, accumulate, regenerate a camera ray
// Every iteration the warp ballots what its lanes want, picks ONE stage (the most populated one) and executes it for
// exactly the lanes that want it; the others keep their state and wait until their stage is picked, by which time
// more lanes have joined them.  Divergent work is thereby re-converged by stage - the in-register equivalent of the
// wavefront queues (generate / traverse / intersect-by-shape-class / shade / env-miss / accumulate) with the ballot
// as the compaction step and no ray state ever leaving the register file.
// ---------------------------------------------------------------------------------------------------------------
enum : uint32_t { W_NODE = 0, W_PQ = 1, W_PF = 2, W_PC = 3, W_SHADE = 4, W_GEN = 5, W_DONE = 6 };
constexpr int kSentinel = 0x7fffffff;
__device__ __forceinline__ uint32_t classOfType(uint32_t type) { return (0x3211211u >> (type * 4u)) & 0xfu; }

template <bool SMEM, bool COUNT, int NODE_STICK>
__global__ void __launch_bounds__(kThreads, 3) traceKernelV2(const RenderParams p)
{
	extern __shared__ __align__(128) float4 smemScene[];
	__shared__ uint64_t mbar;
	SceneView<SMEM> sv;
	if constexpr (SMEM)
	{
		stageSceneToSmem(smemScene, p.scene.sceneBlob, (p.scene.nodeCount + p.scene.primCount) * 64u, &mbar);
		sv.nodes = smemScene;
		sv.prims = smemScene + size_t(p.scene.nodeCount) * 4;
	}
	else
	{
		sv.nodes = p.scene.sceneBlob;
		sv.prims = p.scene.sceneBlob + size_t(p.scene.nodeCount) * 4;
	}

	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t totalPixels = p.width * p.height;
	const V3 camO = mk(p.cam.origin[0], p.cam.origin[1], p.cam.origin[2]);
	constexpr float tMin = 0.001f;

	// path state
	uint32_t want = W_GEN;
	bool hasPath = false, pendingMiss = false;
	uint32_t pixel = kInvalid, sample = p.spp, sampleIdx = 0, bounce = 0, rz = 0, rw = 0;
	V3 color = mk(0.0f, 0.0f, 0.0f);
	V3 ro = camO, rd = mk(0.0f, 0.0f, 1.0f), thr = mk(1.0f, 1.0f, 1.0f), L = mk(0.0f, 0.0f, 0.0f);
	// traversal state
	TravRay tr = makeTravRay(ro, rd);
	float tBest = FLT_MAX;
	int primBest = -1, cur = 0, sp = 0;
	uint32_t sceneBest = 0, leafPrim = 0, leafLeft = 0;
	int stack[kStackSize];
	stack[0] = kSentinel;
	uint32_t rays = 0, nodeVisits = 0, primTests = 0, shades = 0, misses = 0;

	auto setCur = [&](int c)
	{
		cur = c;
		if (c == kSentinel)
		{
			pendingMiss = primBest < 0;
			want = pendingMiss ? W_GEN : W_SHADE;
		}
		else if (c >= 0) want = W_NODE;
		else
		{
			leafPrim = uint32_t(c) & kLeafStartMask;
			leafLeft = (uint32_t(c) >> kLeafCountShift) & 15u;
			want = classOfType((uint32_t(c) >> kLeafTypeShift) & 7u);
			if (leafLeft == 0u) { cur = kSentinel; pendingMiss = primBest < 0; want = pendingMiss ? W_GEN : W_SHADE; } // empty leaf: never referenced by a hit box
		}
	};
	auto pop = [&]() -> int { return sp > 0 ? stack[--sp] : kSentinel; };
	auto startRay = [&]()
	{
		tr = makeTravRay(ro, rd);
		sp = 0; tBest = FLT_MAX; primBest = -1; sceneBest = 0; cur = 0;
		want = W_NODE;
		++rays;
	};
	auto primStep = [&](auto intersect)
	{
		if (COUNT) ++primTests;
		const float4 *pp = sv.prims + leafPrim * 4;
		const float4 r0 = sv.ld(pp), r1 = sv.ld(pp + 1), r2 = sv.ld(pp + 2), meta = sv.ld(pp + 3);
		V3 lo, ld;
		toLocal(r0, r1, r2, ro, rd, lo, ld);
		float t;
		if (intersect(__float_as_uint(meta.x), lo, ld, tBest, t))
		{
			const uint32_t sceneIdx = __float_as_uint(meta.y);
			if (!(t == tBest && primBest >= 0 && sceneIdx < sceneBest))
			{
				tBest = t;
				primBest = int(leafPrim);
				sceneBest = sceneIdx;
			}
		}
		--leafLeft;
		++leafPrim;
		if (leafLeft) want = classOfType(__float_as_uint(sv.ld(sv.prims + leafPrim * 4 + 3).x));
		else setCur(pop());
	};

	while (true)
	{
		// ---- vote: how many lanes want each stage ----
		const uint32_t bN = __ballot_sync(0xffffffffu, want == W_NODE);
		const uint32_t bQ = __ballot_sync(0xffffffffu, want == W_PQ);
		const uint32_t bF = __ballot_sync(0xffffffffu, want == W_PF);
		const uint32_t bC = __ballot_sync(0xffffffffu, want == W_PC);
		const uint32_t bS = __ballot_sync(0xffffffffu, want == W_SHADE);
		const uint32_t bG = __ballot_sync(0xffffffffu, want == W_GEN);
		if ((bN | bQ | bF | bC | bS | bG) == 0u) break;
		uint32_t act = W_NODE;
		int best = __popc(bN);
		{ const int c = __popc(bQ); if (c > best) { best = c; act = W_PQ; } }
		{ const int c = __popc(bF); if (c > best) { best = c; act = W_PF; } }
		{ const int c = __popc(bC); if (c > best) { best = c; act = W_PC; } }
		{ const int c = __popc(bS); if (c > best) { best = c; act = W_SHADE; } }
		{ const int c = __popc(bG); if (c > best) { best = c; act = W_GEN; } }

		if (act == W_NODE)
		{
			// ---- BVH-traverse stage: keep stepping while enough lanes are still walking interior nodes ----
			do
			{
				if (want == W_NODE)
				{
					if (COUNT) ++nodeVisits;
					const float4 *n = sv.nodes + cur * 4;
					const float4 A = sv.ld(n), Bq = sv.ld(n + 1), C = sv.ld(n + 2);
					const float4 Dq = sv.ld(n + 3);
					bool hitA, hitB;
					float nearA, nearB;
					testNodeBoxes(A, Bq, C, tr, tMin, tBest, hitA, hitB, nearA, nearB);
					const int cA = __float_as_int(Dq.x), cB = __float_as_int(Dq.y);
					int next;
					if (hitA && hitB)
					{
						const bool bFirst = nearB < nearA;
						stack[sp++] = bFirst ? cA : cB;
						next = bFirst ? cB : cA;
					}
					else if (hitA) next = cA;
					else if (hitB) next = cB;
					else next = pop();
					setCur(next);
				}
			} while (__popc(__ballot_sync(0xffffffffu, want == W_NODE)) >= NODE_STICK);
		}
		else if (act == W_PQ)
		{
			// ---- intersect stage, quadric class ----
			if (want == W_PQ) primStep([&](uint32_t type, V3 lo, V3 ld, float tMax, float &t) { return intersectQuadric(type, lo, ld, tMin, tMax, t); });
		}
		else if (act == W_PF)
		{
			if (want == W_PF) primStep([&](uint32_t type, V3 lo, V3 ld, float tMax, float &t) { return intersectFlat(type, lo, ld, tMin, tMax, t); });
		}
		else if (act == W_PC)
		{
			if (want == W_PC) primStep([&](uint32_t, V3 lo, V3 ld, float tMax, float &t) { return intersectCube(lo, ld, tMin, tMax, t); });
		}
		else if (act == W_SHADE)
		{
			// ---- shade / sample stage (trace.cu:136-151) ----
			if (want == W_SHADE)
			{
				if (COUNT) ++shades;
				const Surface s = surfaceAt<SMEM>(sv, primBest, ro, rd, tBest);
				const float4 *mp = reinterpret_cast<const float4 *>(p.scene.mats + primBest);
				const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
				L = L + thr * mk(m1.x, m1.y, m1.z);
				V3 base = mk(m0.x, m0.y, m0.z);
				const uint32_t tex = __float_as_uint(m2.x), mtype = __float_as_uint(m2.y);
				if (tex != 0 && tex <= p.scene.texCount)
				{
					const V3 tap = texLookup(p.scene.textures, tex, s.u, s.v);
					base = mk(fastPow(tap.x, 2.2f), fastPow(tap.y, 2.2f), fastPow(tap.z, 2.2f));
				}
				float rnd0, rnd1;
				if (bounce == 0) { rnd0 = uniform01(rz); rnd1 = uniform01(rw); }
				else if (bounce & 1u)
				{
					const uint4 r = philox4x32_10(pixel, sampleIdx, (bounce + 1u) >> 1, 0u, p.seedLo, p.seedHi);
					rnd0 = uniform01(r.x); rnd1 = uniform01(r.y);
					rz = r.z; rw = r.w;
				}
				else { rnd0 = uniform01(rz); rnd1 = uniform01(rw); }
				V3 wi, weight;
				bool cont = sampleMaterial(mtype, base, m0.w, m1.w, s.n, rd, rnd0, rnd1, wi, weight);
				if (cont)
				{
					thr = thr * weight;
					ro = s.p;
					rd = wi;
					++bounce;
					if (bounce >= p.maxBounces) cont = false;
				}
				if (cont) startRay();
				else { want = W_GEN; pendingMiss = false; }
			}
		}
		else
		{
			// ---- env-miss + accumulate + generate stage (trace.cu:115-134, :187-198) ----
			const bool mine = want == W_GEN;
			if (mine && hasPath)
			{
				if (pendingMiss)
				{
					if (COUNT) ++misses;
					if (p.scene.skybox != 0)
					{
						const float theta = acosf(rd.y), phi = atan2f(rd.z, rd.x);
						const V3 sky = texLookup(p.scene.textures, p.scene.skybox, phi / (2.0f * PT_PI), theta / PT_PI);
						L = L + thr * sky;
					}
				}
				color = color + L;
				++sample;
				hasPath = false;
			}
			const bool need = mine && sample == p.spp;
			if (need && pixel != kInvalid)
			{
				float4 out = make_float4(color.x, color.y, color.z, 1.0f);
				if (!p.ignoreHistory)
				{
					const float4 prev = p.accum[pixel];
					out.x += prev.x; out.y += prev.y; out.z += prev.z;
				}
				p.accum[pixel] = out;
			}
			const uint32_t needMask = __ballot_sync(0xffffffffu, need);
			if (needMask)
			{
				const uint32_t leader = __ffs(needMask) - 1;
				unsigned long long base = 0;
				if (lane == leader) base = atomicAdd(&p.counters[kCtrWork], (unsigned long long)__popc(needMask));
				base = __shfl_sync(0xffffffffu, base, leader);
				if (need)
				{
					const unsigned long long m = base + __popc(needMask & ((1u << lane) - 1u));
					if (m >= totalPixels) { want = W_DONE; pixel = kInvalid; }
					else { pixel = uint32_t(m); sample = 0; color = mk(0.0f, 0.0f, 0.0f); }
				}
			}
			if (mine && want == W_GEN)
			{
				sampleIdx = p.sampleOffset + sample * p.sampleStride;
				const uint4 r = philox4x32_10(pixel, sampleIdx, 0u, 0u, p.seedLo, p.seedHi);
				const uint32_t px = pixel % p.width, py = pixel / p.width;
				const float u = divExact(float(px) + uniform01(r.x), float(p.width));
				const float v = divExact(float(py) + uniform01(r.y), float(p.height));
				rz = r.z; rw = r.w;
				ro = camO;
				rd = cameraDir(p.cam, u, v);
				thr = mk(1.0f, 1.0f, 1.0f);
				L = mk(0.0f, 0.0f, 0.0f);
				bounce = 0;
				hasPath = true;
				startRay();
			}
		}
	}

	unsigned long long r64 = rays;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) r64 += __shfl_xor_sync(0xffffffffu, r64, o);
	if (lane == 0) atomicAdd(&p.counters[kCtrRays], r64);
	if (COUNT)
	{
		unsigned long long c[4] = { nodeVisits, primTests, shades, misses };
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
			if (lane == 0) atomicAdd(&p.counters[kCtrNodes + k], c[k]);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// deterministic primary pass / generic ray queries (parity gates) and the output stage
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) primaryKernel(SceneDev scene, CameraDev cam, uint32_t width, uint32_t height, int32_t *hitIndex, float *hitT)
{
	SceneView<false> sv;
	sv.nodes = scene.sceneBlob;
	sv.prims = scene.sceneBlob + size_t(scene.nodeCount) * 4;
	const uint32_t total = width * height;
	uint32_t nv = 0, pt = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
	{
		const uint32_t x = i % width, y = i / width;
		const float u = divExact(float(x) + 0.5f, float(width)), v = divExact(float(y) + 0.5f, float(height));
		const V3 o = mk(cam.origin[0], cam.origin[1], cam.origin[2]);
		const V3 d = cameraDir(cam, u, v);
		const Hit h = closestHit<false, false>(sv, o, d, 0.001f, nv, pt);
		hitIndex[i] = h.prim < 0 ? -1 : int32_t(__float_as_uint(__ldg(sv.prims + h.prim * 4 + 3).y));
		hitT[i] = h.prim < 0 ? 0.0f : h.t;
	}
}

__global__ void __launch_bounds__(kThreads) traceRaysKernel(SceneDev scene, uint32_t n, const float *origins, const float *directions, float tMin,
                                                            int32_t *hitIndex, float *hitT, float *hitNormal)
{
	SceneView<false> sv;
	sv.nodes = scene.sceneBlob;
	sv.prims = scene.sceneBlob + size_t(scene.nodeCount) * 4;
	uint32_t nv = 0, pt = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const V3 o = mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
		const V3 d = mk(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
		const Hit h = closestHit<false, false>(sv, o, d, tMin, nv, pt);
		hitIndex[i] = h.prim < 0 ? -1 : int32_t(__float_as_uint(__ldg(sv.prims + h.prim * 4 + 3).y));
		hitT[i] = h.prim < 0 ? 0.0f : h.t;
		if (hitNormal)
		{
			V3 nn = mk(0.0f, 0.0f, 0.0f);
			if (h.prim >= 0) nn = surfaceAt<false>(sv, h.prim, o, d, h.t).n;
			hitNormal[3 * i] = nn.x; hitNormal[3 * i + 1] = nn.y; hitNormal[3 * i + 2] = nn.z;
		}
	}
}

// tonemap (kernels/tonemap.cu:4-27): mean -> Reinhard -> gamma 1/2.2 -> truncating RGBA8.  HBM-bound: 16 B in, 4 B out.
__global__ void __launch_bounds__(kThreads) tonemapKernel(const float4 *__restrict__ accum, uchar4 *__restrict__ out, uint32_t pixels, float invSampleCount)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x)
	{
		const float4 a = __ldg(accum + i);
		float r = a.x * invSampleCount, g = a.y * invSampleCount, b = a.z * invSampleCount;
		r = r * (1.0f / (r + 1.0f)); g = g * (1.0f / (g + 1.0f)); b = b * (1.0f / (b + 1.0f));
		r = powf(r, 1.0f / 2.2f); g = powf(g, 1.0f / 2.2f); b = powf(b, 1.0f / 2.2f);
		out[i] = make_uchar4((unsigned char)(r * 255.0f), (unsigned char)(g * 255.0f), (unsigned char)(b * 255.0f), 255);
	}
}

// getHDRImageData's host loop (Pathtracer.cpp:307-312) on the device: all four channels scaled
__global__ void __launch_bounds__(kThreads) scaleKernel(const float4 *__restrict__ accum, float4 *__restrict__ out, uint32_t pixels, float scale)
{
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += gridDim.x * blockDim.x)
	{
		const float4 a = __ldg(accum + i);
		out[i] = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
template <typename K>
static int launchKernel(K kern, const RenderParams &p, const LaunchConfig &cfg, size_t smemBytes, cudaStream_t stream)
{
	if (smemBytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemBytes));
	int blocksPerSm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, kern, kThreads, smemBytes);
	if (blocksPerSm < 1) blocksPerSm = 1;
	const int grid = cfg.smCount * blocksPerSm;
	kern<<<grid, kThreads, smemBytes, stream>>>(p);
	return 1;
}

int launchTrace(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem)
{
	const size_t sceneBytes = (size_t(p.scene.nodeCount) + p.scene.primCount) * 64;
	// leave room for 2+ CTAs per SM when the scene is small; a scene larger than the opt-in limit stays in L2/HBM
	const bool smem = cfg.smemScene && sceneBytes + 1024 <= cfg.maxSmemOptin;
	if (usedSmem) *usedSmem = smem ? 1 : 0;
	const size_t sb = smem ? sceneBytes : 0;
#define PT_PICK(KERN, ...)                                                                                                        \
	(smem ? (cfg.countWork ? launchKernel(KERN<true, true __VA_ARGS__>, p, cfg, sb, stream) : launchKernel(KERN<true, false __VA_ARGS__>, p, cfg, sb, stream)) \
	      : (cfg.countWork ? launchKernel(KERN<false, true __VA_ARGS__>, p, cfg, sb, stream) : launchKernel(KERN<false, false __VA_ARGS__>, p, cfg, sb, stream)))
	switch (cfg.variant)
	{
	case 1: return PT_PICK(traceKernel, , 0);      // per-lane if/else traversal
	case 4: return PT_PICK(traceKernel, , 1);      // while-while traversal
	case 2: return PT_PICK(traceKernelV2, , 8);    // stage-voting warp scheduler
	case 3: return PT_PICK(traceKernelV2, , 16);
	default: return PT_PICK(traceKernel, , 2);     // while-while + speculative leaf parking (fastest measured)
	}
#undef PT_PICK
}

int launchPrimary(const SceneDev &scene, const CameraDev &cam, uint32_t width, uint32_t height, int32_t *hitIndex, float *hitT, cudaStream_t stream)
{
	const uint32_t total = width * height;
	const int grid = int((total + kThreads - 1) / kThreads);
	primaryKernel<<<grid, kThreads, 0, stream>>>(scene, cam, width, height, hitIndex, hitT);
	return 1;
}

int launchTraceRays(const SceneDev &scene, size_t n, const float *origins, const float *directions, float tMin, int32_t *hitIndex, float *hitT,
                    float *hitNormal, cudaStream_t stream)
{
	const int grid = int((n + kThreads - 1) / kThreads);
	traceRaysKernel<<<grid, kThreads, 0, stream>>>(scene, uint32_t(n), origins, directions, tMin, hitIndex, hitT, hitNormal);
	return 1;
}

int launchTonemap(const float4 *accum, uchar4 *out, uint32_t pixels, float invSampleCount, cudaStream_t stream)
{
	const int grid = int(min((pixels + kThreads - 1) / kThreads, 148u * 8u));
	tonemapKernel<<<grid, kThreads, 0, stream>>>(accum, out, pixels, invSampleCount);
	return 1;
}

int launchScale(const float4 *accum, float4 *out, uint32_t pixels, float scale, cudaStream_t stream)
{
	const int grid = int(min((pixels + kThreads - 1) / kThreads, 148u * 8u));
	scaleKernel<<<grid, kThreads, 0, stream>>>(accum, out, pixels, scale);
	return 1;
}

} // namespace ptb
