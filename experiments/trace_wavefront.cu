// trace_wavefront.cu — the persistent-thread WAVEFRONT pipeline of the trace path (default kernel).
//
// One persistent CTA per SM (24 warps).  The CTA owns a POOL of path slots in shared memory (struct-of-arrays, 23 words
// per slot: ray, throughput, path radiance, pixel colour, pixel, sample/bounce state, leftover Philox words, closest hit,
// two deferred leaves) - about 1.7x as many paths as the CTA has lanes - and four multi-producer/multi-consumer QUEUES
// of slot ids, also in shared memory, one per pipeline stage:
//
//     gen    path ended (env miss, absorbed, max depth) or slot is new:  env-miss lookup (trace.cu:115-134), accumulate
//            (:196-198), next sample / next pixel (warp-aggregated global atomic), generate the camera ray (:187-192),
//            test the scene-spanning primitives                                       -> ready (or the slot retires)
//     ready  has a ray: BVH traversal (hitBVH, trace.cu:28-98).  The ray lives in the lane's registers; leaves met on
//            the way are DEFERRED (two per ray), so the node loop is the only code in the loop      -> leaf | hit | gen
//     leaf   deferred primitive tests (Hittable::hit, Hittable.inl:88-145)                              -> hit | gen
//     hit    shade / sample (trace.cu:136-151), then the scene-spanning primitives for the new ray  -> ready | gen
//
// Every warp repeatedly picks the stage that can fill most of its lanes, pops up to 32 slot ids from that queue
// (one atomicCAS + one atomicAdd per warp, ballot/popc prefix for the lanes), runs the stage and pushes the slots to
// their next queues.  Whatever mixture of path depths, materials and hit/miss outcomes the pool holds, each stage
// therefore executes with full warps (the queues ARE the compaction), and a lane whose traversal retires is refilled
// from `ready` at the next round: the node loop leaves when fewer than `nodeLow` lanes are still walking instead of
// waiting for the longest ray of the warp.  A slot owns its pixel for all of its samples, added in order, so the image
// does not depend on scheduling: bit-identical to the per-lane kernels in trace_kernels.cu.
#include "trace_common.cuh"
#include <algorithm>

namespace ptb
{

constexpr int kWfThreads = 768;
constexpr int kWfWarps = kWfThreads / 32;
enum : int { S_OX = 0, S_OY, S_OZ, S_DX, S_DY, S_DZ, S_TX, S_TY, S_TZ, S_LX, S_LY, S_LZ, S_CX, S_CY, S_CZ, S_PIXEL, S_STATE, S_RZ, S_RW, S_T, S_PRIM, S_PARK0, S_PARK1, kWfWords };
enum : int { Q_READY = 0, Q_DONE = 1, kWfQueues };
enum : int { L_LEAF = 0, L_HIT, L_GEN, L_OUT, kWfLists };
constexpr int kWfListCap = 128;
constexpr uint32_t kWfSampleMask = 0x00ffffffu, kWfBounceShift = 24, kWfHasPath = 0x80000000u;
constexpr uint32_t kQEmpty = 0xffffu;

struct WfCtrl
{
	uint32_t cnt[kWfQueues];  // committed entries per queue
	uint32_t head[kWfQueues]; // next position to consume
	uint32_t tail[kWfQueues]; // next position to produce
	uint32_t live;            // slots not yet retired
	uint32_t error;           // watchdog
};

struct WfLayout
{
	uint32_t sceneBytes; // aligned
	uint32_t slots;      // N
	uint32_t ring;       // power of two >= N
	uint32_t warps, traceWarps;
	size_t total() const
	{
		return size_t(sceneBytes) + size_t(kWfWords) * slots * 4 + size_t(kWfQueues) * ring * 2 + size_t(warps - traceWarps) * kWfLists * kWfListCap * 2 + sizeof(WfCtrl);
	}
};

// closest-hit fold with the tie rule (equal t: larger scene index wins, Q7); the scene index of the current best is only
// fetched when a tie actually happens.  One out-of-line copy: the pipeline calls it from four places and the kernel has
// to stay inside the instruction cache (warps of one SM are in different stages at the same time).
#ifndef WF_FOLD_INLINE
#define WF_FOLD_INLINE __noinline__
#endif
struct WfBest
{
	float t;
	int prim;
};
template <bool SMEM>
__device__ WF_FOLD_INLINE WfBest wfFoldPrim(const float4 *prims, uint32_t prim, V3 o, V3 d, WfBest best)
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
__global__ void __launch_bounds__(kWfThreads, 1) traceKernelWF(const RenderParams p, const WfLayout lay, const int nodeLow, const int readyLow)
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
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;
	const uint32_t N = lay.slots, ringMask = lay.ring - 1u;
	float *pool = reinterpret_cast<float *>(reinterpret_cast<char *>(smemScene) + lay.sceneBytes);
	uint32_t *poolU = reinterpret_cast<uint32_t *>(pool);
	volatile uint16_t *ring = reinterpret_cast<volatile uint16_t *>(pool + size_t(kWfWords) * N);
	uint16_t *lists = const_cast<uint16_t *>(ring) + size_t(kWfQueues) * lay.ring;
	WfCtrl *ctrl = reinterpret_cast<WfCtrl *>(lists + size_t(lay.warps - lay.traceWarps) * kWfLists * kWfListCap);
	volatile WfCtrl *vctrl = ctrl;
#define SF(field, slot) pool[(field) * N + (slot)]
#define SU(field, slot) poolU[(field) * N + (slot)]

	// ---- pool / queue initialisation: every slot starts in the DONE queue as a finished "path" without a pixel ----
	for (uint32_t i = threadIdx.x; i < kWfQueues * lay.ring; i += blockDim.x) ring[i] = uint16_t(kQEmpty);
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
	{
		SU(S_PIXEL, i) = kInvalid;
		SU(S_STATE, i) = p.spp & kWfSampleMask;
		SF(S_CX, i) = 0.0f; SF(S_CY, i) = 0.0f; SF(S_CZ, i) = 0.0f;
		SU(S_PRIM, i) = 0xffffffffu; SU(S_PARK0, i) = uint32_t(kEmptyChild); SU(S_PARK1, i) = uint32_t(kEmptyChild);
		ring[Q_DONE * lay.ring + i] = uint16_t(i);
	}
	if (threadIdx.x == 0)
	{
		for (int q = 0; q < kWfQueues; ++q) { ctrl->cnt[q] = 0; ctrl->head[q] = 0; ctrl->tail[q] = 0; }
		ctrl->cnt[Q_DONE] = N; ctrl->tail[Q_DONE] = N;
		ctrl->live = N;
		ctrl->error = 0;
	}
	__syncthreads();

	const uint32_t totalPixels = p.width * p.height;
	const float invW = 1.0f / float(p.width), invH = 1.0f / float(p.height);
	const V3 camO = mk(p.cam.origin[0], p.cam.origin[1], p.cam.origin[2]);
	constexpr float tMin = 0.001f;
	uint32_t rays = 0, nodeVisits = 0, primTests = 0, shades = 0, misses = 0;
	uint32_t dbg[13] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }; // lane 0: kCtrTraceRounds.. (count_work only)
#define DBG(idx, v) do { if (COUNT && lane == 0) dbg[(idx) - kCtrTraceRounds] += (v); } while (0)

	// ---- queue operations (called by all 32 lanes of the warp) ----
	// pop up to `want` slot ids; returns how many, lane i < n gets the i-th in `item`
	auto qPop = [&](int q, uint32_t want, uint32_t &item) -> uint32_t
	{
		uint32_t n = 0, h = 0;
		if (lane == 0)
		{
			uint32_t c = vctrl->cnt[q];
			while (c > 0)
			{
				const uint32_t take = min(c, want);
				const uint32_t old = atomicCAS(&ctrl->cnt[q], c, c - take);
				if (old == c) { n = take; break; }
				c = old;
			}
			if (n) h = atomicAdd(&ctrl->head[q], n);
		}
		n = __shfl_sync(full, n, 0);
		h = __shfl_sync(full, h, 0);
		item = 0;
		if (lane < n)
		{
			volatile uint16_t *e = ring + q * lay.ring + ((h + lane) & ringMask);
			uint32_t v = *e;
			for (uint32_t spin = 0; v == kQEmpty; ++spin)
			{
				if (spin > (1u << 22)) { ctrl->error = 2; v = 0; break; }
				v = *e;
			}
			*e = uint16_t(kQEmpty);
			if (v >= N) { ctrl->error = 3; v = 0; }
			item = v;
		}
		__threadfence_block(); // acquire: the slot's words were written before its id was published
		return n;
	};
	// push the slots of the lanes with pred set
	auto qPush = [&](int q, bool pred, uint32_t s)
	{
		const uint32_t m = __ballot_sync(full, pred);
		if (m == 0u) return;
		const uint32_t n = __popc(m), leader = __ffs(m) - 1;
		__threadfence_block(); // release: slot words before the id
		uint32_t base = 0;
		if (lane == leader) base = atomicAdd(&ctrl->tail[q], n);
		base = __shfl_sync(full, base, leader);
		if (pred)
		{
			volatile uint16_t *e = ring + q * lay.ring + ((base + __popc(m & ltMask)) & ringMask);
			if (*e != kQEmpty) ctrl->error = 6;
			if (s >= N) ctrl->error = 7;
			*e = uint16_t(s);
		}
		__threadfence_block();
		__syncwarp();
		if (lane == leader) atomicAdd(&ctrl->cnt[q], n);
	};
	// bit 31: watchdog fired, bit 30: every slot has retired, bits 0..13: committed entries of queue q, bits 14..27: of READY
	auto readCtl = [&](int q) -> uint32_t
	{
		__syncwarp();
		uint32_t w = 0;
		if (lane == 0) w = (vctrl->error != 0u ? 0x80000000u : 0u) | (vctrl->live == 0u ? 0x40000000u : 0u) | min(vctrl->cnt[q], 0x3fffu) | (min(vctrl->cnt[Q_READY], 0x3fffu) << 14);
		return __shfl_sync(full, w, 0);
	};
	auto foldLeaf = [&](int leaf, V3 o, V3 d, float &tBest, int &primBest)
	{
		const uint32_t first = uint32_t(leaf) & kLeafStartMask;
		const uint32_t count = (uint32_t(leaf) >> kLeafCountShift) & 15u;
		for (uint32_t i = 0; i < count; ++i)
		{
			if (COUNT) ++primTests;
			WfBest b;
			b.t = tBest; b.prim = primBest;
			if (first + i >= p.scene.primCount) { ctrl->error = 13; break; }
			b = wfFoldPrim<SMEM>(sv.prims, first + i, o, d, b);
			tBest = b.t; primBest = b.prim;
		}
	};

	if (warp < lay.traceWarps)
	{
		// =====================================================================================================
		// TRACE warps: BVH traversal only (hitBVH, trace.cu:28-98).  READY -> registers -> DONE.
		// =====================================================================================================
		int slot = -1;
		V3 ro = camO, rd = mk(0.0f, 0.0f, 1.0f);
		TravRay tr = makeTravRay(ro, rd);
		int cur = kEmptyChild, park0 = kEmptyChild, park1 = kEmptyChild, sp = 0, primBest = -1;
		float tBest = FLT_MAX;
		int stack[kStackSize];
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
		for (uint32_t iter = 0;; ++iter)
		{
			if (iter > (1u << 26)) ctrl->error = 1; // watchdog: never hang the GPU on a scheduling bug
			// every scheduling decision must be warp-uniform: ONE lane reads the shared control words and broadcasts them
			// (lanes that woke up from __nanosleep at different times would otherwise read different values and part ways)
			const uint32_t ctl = readCtl(Q_READY);
			if (ctl & 0x80000000u) break;
			const uint32_t inflight = __popc(__ballot_sync(full, slot >= 0));
			if (inflight < 32u && (ctl & 0x3fffu) > 0u)
			{
				uint32_t item;
				const uint32_t n = qPop(Q_READY, 32u - inflight, item);
				DBG(kCtrRefills, 1); DBG(kCtrRefillSlots, n);
				const uint32_t needRay = __ballot_sync(full, slot < 0);
				const uint32_t r = __popc(needRay & ltMask);
				const uint32_t got = __shfl_sync(full, item, r & 31u);
				if (slot < 0 && r < n)
				{
					slot = int(got);
					if (uint32_t(slot) >= N) { ctrl->error = 8; slot = 0; }
					ro = mk(SF(S_OX, slot), SF(S_OY, slot), SF(S_OZ, slot));
					rd = mk(SF(S_DX, slot), SF(S_DY, slot), SF(S_DZ, slot));
					tr = makeTravRay(ro, rd);
					tBest = SF(S_T, slot);
					primBest = int(SU(S_PRIM, slot));
					stack[0] = kEmptyChild; // sentinel: a leaf reference with zero primitives
					sp = 1; cur = 0; park0 = kEmptyChild; park1 = kEmptyChild;
					++rays;
				}
			}
			else if (inflight == 0u)
			{
				if (ctl & 0x40000000u) break;
				DBG(kCtrIdle, 1);
				__nanosleep(100);
				continue;
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
					if (uint32_t(cur) >= p.scene.nodeCount) { ctrl->error = 11; cur = 0; }
					if (sp < 1 || sp >= kStackSize - 1) { ctrl->error = 12; sp = 1; }
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
				if (done)
				{
					SF(S_T, slot) = tBest; SU(S_PRIM, slot) = uint32_t(primBest);
					SU(S_PARK0, slot) = uint32_t(park0); SU(S_PARK1, slot) = uint32_t(park1);
				}
				qPush(Q_DONE, done, uint32_t(slot));
				if (done) slot = -1;
			}
		}
	}
	else
	{
		// =====================================================================================================
		// STAGE warps: DONE -> classify into warp-local lists (leaf / hit / gen) -> run a stage whenever a list holds a
		// full warp -> READY.  Only the two global queues use atomics; the lists are private to the warp.
		// =====================================================================================================
		uint16_t *myLists = lists + size_t(warp - lay.traceWarps) * kWfLists * kWfListCap;
		uint32_t cntL[kWfLists] = { 0, 0, 0, 0 }; // warp-uniform
		auto listAppend = [&](int l, bool pred, uint32_t s)
		{
			const uint32_t m = __ballot_sync(full, pred);
			if (pred) myLists[l * kWfListCap + cntL[l] + __popc(m & ltMask)] = uint16_t(s);
			cntL[l] += __popc(m);
			if (cntL[l] > uint32_t(kWfListCap)) ctrl->error = 4;
		};
		auto listTake = [&](int l, uint32_t &s) -> uint32_t
		{
			__syncwarp();
			const uint32_t n = min(cntL[l], 32u);
			cntL[l] -= n;
			s = lane < n ? uint32_t(myLists[l * kWfListCap + cntL[l] + lane]) : 0u;
			if (s >= N) { ctrl->error = 9; s = 0; }
			__syncwarp();
			return n;
		};
		for (uint32_t iter = 0;; ++iter)
		{
			if (iter > (1u << 26)) ctrl->error = 1;
			const uint32_t ctl = readCtl(Q_DONE);
			if (ctl & 0x80000000u) break;
			int stage = -1; // 0 leaf, 1 shade, 2 gen, 3 flush READY, 4 pop DONE
			const uint32_t cDone = ctl & 0x3fffu, cReady = (ctl >> 14) & 0x3fffu;
			// full lists first, downstream first (bounds every list by 63 entries)
			if (cntL[L_OUT] >= 32u) stage = 3;
			else if (cntL[L_GEN] >= 32u) stage = 2;
			else if (cntL[L_HIT] >= 32u) stage = 1;
			else if (cntL[L_LEAF] >= 32u) stage = 0;
			else if (cReady < uint32_t(readyLow) && (cntL[L_OUT] | cntL[L_GEN] | cntL[L_HIT] | cntL[L_LEAF]) != 0u)
			{
				// the trace warps are about to starve: feed them from partial lists, shortest way to READY first
				if (cntL[L_OUT] > 0u) stage = 3;
				else if (cntL[L_GEN] >= cntL[L_HIT] && cntL[L_GEN] >= cntL[L_LEAF]) stage = 2;
				else if (cntL[L_HIT] >= cntL[L_LEAF]) stage = 1;
				else stage = 0;
			}
			else if (cDone > 0u) stage = 4;
			else if (cntL[L_OUT] > 0u) stage = 3;
			else if (cntL[L_LEAF] > 0u) stage = 0; // nothing arrives any more: drain the partial lists, upstream first
			else if (cntL[L_HIT] > 0u) stage = 1;
			else if (cntL[L_GEN] > 0u) stage = 2;
			if (stage < 0)
			{
				if (ctl & 0x40000000u) break;
				DBG(kCtrIdle, 1);
				__nanosleep(100);
				continue;
			}

			if (stage == 4)
			{
				uint32_t s;
				const uint32_t n = qPop(Q_DONE, 32u, s);
				const bool mine = lane < n;
				const bool needLeaf = mine && int(SU(S_PARK0, s)) != kEmptyChild;
				const bool isHit = mine && !needLeaf && int(SU(S_PRIM, s)) >= 0;
				listAppend(L_LEAF, needLeaf, s);
				listAppend(L_HIT, isHit, s);
				listAppend(L_GEN, mine && !needLeaf && !isHit, s);
				continue;
			}
			if (stage == 3)
			{
				uint32_t s;
				const uint32_t n = listTake(L_OUT, s);
				qPush(Q_READY, lane < n, s);
				continue;
			}
			if (stage == 0)
			{
				// ---------------- deferred primitive tests (Hittable::hit, Hittable.inl:88-145) ----------------
				uint32_t s;
				const uint32_t n = listTake(L_LEAF, s);
				DBG(kCtrLeafExec, 1); DBG(kCtrLeafSlots, n);
				const bool mine = lane < n;
				int lprim = -1;
				if (mine)
				{
					const V3 lro = mk(SF(S_OX, s), SF(S_OY, s), SF(S_OZ, s)), lrd = mk(SF(S_DX, s), SF(S_DY, s), SF(S_DZ, s));
					float lt = SF(S_T, s);
					lprim = int(SU(S_PRIM, s));
					const int l0 = int(SU(S_PARK0, s)), l1 = int(SU(S_PARK1, s));
					foldLeaf(l0, lro, lrd, lt, lprim);
					foldLeaf(l1, lro, lrd, lt, lprim); // kEmptyChild has count 0
					SF(S_T, s) = lt; SU(S_PRIM, s) = uint32_t(lprim);
				}
				listAppend(L_HIT, mine && lprim >= 0, s);
				listAppend(L_GEN, mine && lprim < 0, s);
				continue;
			}
			if (stage == 1)
			{
				// ---------------- shade / sample stage (trace.cu:136-151) ----------------
				uint32_t s;
				const uint32_t n = listTake(L_HIT, s);
				DBG(kCtrShadeExec, 1); DBG(kCtrShadeSlots, n);
				const bool mine = lane < n;
				bool cont = false;
				if (mine)
				{
					if (COUNT) ++shades;
					const V3 sro = mk(SF(S_OX, s), SF(S_OY, s), SF(S_OZ, s)), srd = mk(SF(S_DX, s), SF(S_DY, s), SF(S_DZ, s));
					int prim = int(SU(S_PRIM, s));
					if (uint32_t(prim) >= p.scene.primCount) { ctrl->error = 5; prim = 0; }
					const Surface sf = surfaceAt<SMEM>(sv, prim, sro, srd, SF(S_T, s));
					const float4 *mp = reinterpret_cast<const float4 *>(p.scene.mats + prim);
					const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
					V3 thr = mk(SF(S_TX, s), SF(S_TY, s), SF(S_TZ, s));
					const V3 L = mk(SF(S_LX, s), SF(S_LY, s), SF(S_LZ, s)) + thr * mk(m1.x, m1.y, m1.z); // getEmitted, Material.inl:62-65
					SF(S_LX, s) = L.x; SF(S_LY, s) = L.y; SF(S_LZ, s) = L.z;
					V3 base = mk(m0.x, m0.y, m0.z);
					const uint32_t tex = __float_as_uint(m2.x), mtype = __float_as_uint(m2.y);
					if (tex != 0 && tex <= p.scene.texCount)
					{
						const V3 tap = texLookup(p.scene.textures, tex, sf.u, sf.v); // Material.inl:26-35
						base = mk(fastPow(tap.x, 2.2f), fastPow(tap.y, 2.2f), fastPow(tap.z, 2.2f));
					}
					const uint32_t state = SU(S_STATE, s);
					uint32_t bounce = (state >> kWfBounceShift) & 0x7fu;
					float rnd0, rnd1;
					if (bounce != 0u && (bounce & 1u))
					{
						const uint4 r = philox4x32_10(SU(S_PIXEL, s), p.sampleOffset + (state & kWfSampleMask) * p.sampleStride, (bounce + 1u) >> 1, 0u, p.seedLo, p.seedHi);
						rnd0 = uniform01(r.x); rnd1 = uniform01(r.y);
						SU(S_RZ, s) = r.z; SU(S_RW, s) = r.w;
					}
					else { rnd0 = uniform01(SU(S_RZ, s)); rnd1 = uniform01(SU(S_RW, s)); }
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
						SF(S_TX, s) = thr.x; SF(S_TY, s) = thr.y; SF(S_TZ, s) = thr.z;
						SF(S_OX, s) = sf.p.x; SF(S_OY, s) = sf.p.y; SF(S_OZ, s) = sf.p.z;
						SF(S_DX, s) = wi.x; SF(S_DY, s) = wi.y; SF(S_DZ, s) = wi.z;
						SU(S_STATE, s) = (state & ~(0x7fu << kWfBounceShift)) | (bounce << kWfBounceShift);
						// the scene-spanning primitives, tested here where the warp is full and every lane reads the same record
						WfBest gb;
						gb.t = FLT_MAX; gb.prim = -1;
						for (uint32_t g = 0; g < sv.globalCount; ++g)
						{
							if (COUNT) ++primTests;
							gb = wfFoldPrim<SMEM>(sv.prims, g, sf.p, wi, gb);
						}
						SF(S_T, s) = gb.t; SU(S_PRIM, s) = uint32_t(gb.prim);
					}
				}
				listAppend(L_OUT, cont, s);
				listAppend(L_GEN, mine && !cont, s); // S_PRIM stays >= 0: the generate stage will not look up the environment
				continue;
			}
			// ---------------- env-miss + accumulate + generate stage (trace.cu:115-134, :187-198) ----------------
			{
				uint32_t s;
				const uint32_t n = listTake(L_GEN, s);
				DBG(kCtrGenExec, 1); DBG(kCtrGenSlots, n);
				const bool mine = lane < n;
				uint32_t pixel = kInvalid, sample = 0;
				V3 color = mk(0.0f, 0.0f, 0.0f);
				bool need = false;
				if (mine)
				{
					const uint32_t state = SU(S_STATE, s);
					pixel = SU(S_PIXEL, s);
					if (pixel != kInvalid && pixel >= totalPixels) { ctrl->error = 10; pixel = kInvalid; }
					sample = state & kWfSampleMask;
					color = mk(SF(S_CX, s), SF(S_CY, s), SF(S_CZ, s));
					if (state & kWfHasPath)
					{
						V3 L = mk(SF(S_LX, s), SF(S_LY, s), SF(S_LZ, s));
						if (int(SU(S_PRIM, s)) < 0)
						{
							if (COUNT) ++misses;
							if (p.scene.skybox != 0)
							{
								const V3 mrd = mk(SF(S_DX, s), SF(S_DY, s), SF(S_DZ, s)), thr = mk(SF(S_TX, s), SF(S_TY, s), SF(S_TZ, s));
								const float theta = fastAcos(mrd.y), phi = fastAtan2(mrd.z, mrd.x);
								const V3 sky = texLookup(p.scene.textures, p.scene.skybox, phi * (0.5f / PT_PI), theta * (1.0f / PT_PI));
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
						if (m >= totalPixels) alive = false; // the slot retires
						else { pixel = uint32_t(m); sample = 0; color = mk(0.0f, 0.0f, 0.0f); }
					}
					const uint32_t deadMask = __ballot_sync(full, mine && !alive);
					if (deadMask != 0u && lane == 0) atomicSub(&ctrl->live, uint32_t(__popc(deadMask)));
				}
				if (alive)
				{
					const uint32_t sampleIdx = p.sampleOffset + sample * p.sampleStride;
					const uint4 r = philox4x32_10(pixel, sampleIdx, 0u, 0u, p.seedLo, p.seedHi);
					uint32_t px, py;
				pixelToXY(pixel, p.width, p.height, px, py);
					const float u = (float(px) + uniform01(r.x)) * invW; // trace.cu:190
					const float v = (float(py) + uniform01(r.y)) * invH;
					const V3 d = cameraDir<kHotExact>(p.cam, u, v);
					SF(S_OX, s) = camO.x; SF(S_OY, s) = camO.y; SF(S_OZ, s) = camO.z;
					SF(S_DX, s) = d.x; SF(S_DY, s) = d.y; SF(S_DZ, s) = d.z;
					SF(S_TX, s) = 1.0f; SF(S_TY, s) = 1.0f; SF(S_TZ, s) = 1.0f;
					SF(S_LX, s) = 0.0f; SF(S_LY, s) = 0.0f; SF(S_LZ, s) = 0.0f;
					SF(S_CX, s) = color.x; SF(S_CY, s) = color.y; SF(S_CZ, s) = color.z;
					SU(S_RZ, s) = r.z; SU(S_RW, s) = r.w;
					SU(S_PIXEL, s) = pixel;
					SU(S_STATE, s) = kWfHasPath | sample;
					WfBest gb;
					gb.t = FLT_MAX; gb.prim = -1;
					for (uint32_t g = 0; g < sv.globalCount; ++g)
					{
						if (COUNT) ++primTests;
						gb = wfFoldPrim<SMEM>(sv.prims, g, camO, d, gb);
					}
					SF(S_T, s) = gb.t; SU(S_PRIM, s) = uint32_t(gb.prim);
				}
				listAppend(L_OUT, alive, s);
			}
		}
	}
#undef SF
#undef SU

	if (vctrl->error != 0u && threadIdx.x == 0) atomicMax(&p.counters[kCtrError], (unsigned long long)vctrl->error);
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
static int launchWf(K kern, const RenderParams &p, const LaunchConfig &cfg, const WfLayout &lay, cudaStream_t stream)
{
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lay.total()));
	const int nodeLow = cfg.nodeLow > 0 ? cfg.nodeLow : 24;
	kern<<<cfg.smCount, lay.warps * 32, lay.total(), stream>>>(p, lay, nodeLow, cfg.readyLow >= 0 ? cfg.readyLow : 128);
	return 1;
}

// Returns 0 when the wavefront kernel cannot run this configuration (the caller falls back to the per-lane kernel).
int launchTraceWavefront(const RenderParams &p, const LaunchConfig &cfg, cudaStream_t stream, int *usedSmem)
{
	if (p.spp > kWfSampleMask || p.maxBounces > 0x7fu) return 0;
	const size_t sceneBytes = (size_t(p.scene.nodeCount) + p.scene.primCount) * 64;
	const size_t aligned = (sceneBytes + 127) & ~size_t(127);
	const size_t avail = cfg.maxSmemOptin > 1024 ? cfg.maxSmemOptin - 1024 : 0;
	WfLayout lay;
	lay.warps = uint32_t(cfg.poolWarps > 0 ? std::min(cfg.poolWarps, kWfWarps) : kWfWarps);
	lay.traceWarps = uint32_t(cfg.traceWarps > 0 ? std::min<int>(cfg.traceWarps, int(lay.warps) - 1) : int(lay.warps) / 2);
	lay.slots = cfg.poolSlots > 0 ? uint32_t(cfg.poolSlots) : 1280u;
	lay.ring = 1;
	while (lay.ring < lay.slots) lay.ring <<= 1;
	lay.sceneBytes = uint32_t(aligned);
	bool smem = cfg.smemScene && lay.total() <= avail;
	if (!smem)
	{
		lay.sceneBytes = 0;
		if (cfg.poolSlots <= 0) lay.slots = 1536u;
		lay.ring = 1;
		while (lay.ring < lay.slots) lay.ring <<= 1;
		if (lay.total() > avail) return 0;
	}
	if (lay.slots >= kQEmpty) return 0;
	if (usedSmem) *usedSmem = smem ? 1 : 0;
	if (smem) return cfg.countWork ? launchWf(traceKernelWF<true, true>, p, cfg, lay, stream) : launchWf(traceKernelWF<true, false>, p, cfg, lay, stream);
	return cfg.countWork ? launchWf(traceKernelWF<false, true>, p, cfg, lay, stream) : launchWf(traceKernelWF<false, false>, p, cfg, lay, stream);
}

} // namespace ptb
