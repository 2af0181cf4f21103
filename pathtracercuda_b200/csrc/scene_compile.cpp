// scene_compile.cpp — host scene compiler (see scene_compile.h).
//
// What it replaces in the reference: CpuHittable's ctor (Hittable.cpp:115-179), BVH::build (BVH.cpp:66-228) and the
// Hittable/BVHNode upload format (Pathtracer.cpp:125-155).  The transform arithmetic keeps the reference's float
// operation order (this file is compiled with -ffp-contract=off) because the world->local rows feed straight into
// the hit distance t that the parity gate compares; everything else (tree builder, node format) is our own design:
// 16-bin SAH over centroid bounds, leaves chosen by cost, two-box 64-byte nodes, separate material table.
#include "scene_compile.h"
#include <algorithm>
#include <atomic>
#ifdef _OPENMP
#include <parallel/algorithm>
#endif
#include <cfloat>
#include <cmath>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace ptb
{

namespace
{
struct V3 { float x, y, z; };
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

void quatToRotMat(const float q[4], float m[3][3])
{
	const float x = q[0], y = q[1], z = q[2], w = q[3];
	const float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
	m[0][0] = 1.0f - 2.0f * (yy + zz); m[0][1] = 2.0f * (xy + wz); m[0][2] = 2.0f * (xz - wy);
	m[1][0] = 2.0f * (xy - wz); m[1][1] = 1.0f - 2.0f * (xx + zz); m[1][2] = 2.0f * (yz + wx);
	m[2][0] = 2.0f * (xz + wy); m[2][1] = 2.0f * (yz - wx); m[2][2] = 1.0f - 2.0f * (xx + yy);
}
} // namespace

void computeObjectXform(const pt_object_desc &d, ObjectXform &o)
{
	V3 scale = { d.scale[0], d.scale[1], d.scale[2] };
	if (d.type == PT_DISK || d.type == PT_QUAD)
	{
		scale.y = 1.0f; // flat shapes ignore their y scale (Hittable.cpp:124-128)
	}
	const V3 pos = { d.position[0], d.position[1], d.position[2] };

	// Euler XYZ (radians) -> quaternion -> rotation; inverse via the conjugate (Hittable.cpp:33-56)
	const float hx = d.rotation[0] * 0.5f, hy = d.rotation[1] * 0.5f, hz = d.rotation[2] * 0.5f;
	const float cx = cosf(hx), cy = cosf(hy), cz = cosf(hz), sx = sinf(hx), sy = sinf(hy), sz = sinf(hz);
	float q[4], qi[4];
	q[3] = cx * cy * cz + sx * sy * sz;
	q[0] = sx * cy * cz - cx * sy * sz;
	q[1] = cx * sy * cz + sx * cy * sz;
	q[2] = cx * cy * sz - sx * sy * cz;
	const float invDot = 1.0f / (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
	qi[0] = -q[0] * invDot; qi[1] = -q[1] * invDot; qi[2] = -q[2] * invDot; qi[3] = q[3] * invDot;

	float rInv[3][3], r[3][3];
	quatToRotMat(qi, rInv);
	quatToRotMat(q, r);
	const float invS[3] = { 1.0f / scale.x, 1.0f / scale.y, 1.0f / scale.z };
	const V3 npos = { -pos.x, -pos.y, -pos.z };
	for (int k = 0; k < 3; ++k) // world->local = S^-1 R^-1 T^-1 (Hittable.cpp:59-79)
	{
		o.w2l[k][0] = invS[k] * rInv[0][k];
		o.w2l[k][1] = invS[k] * rInv[1][k];
		o.w2l[k][2] = invS[k] * rInv[2][k];
		o.w2l[k][3] = invS[k] * dot3(V3{ rInv[0][k], rInv[1][k], rInv[2][k] }, npos);
	}
	const float p[3] = { pos.x, pos.y, pos.z };
	for (int k = 0; k < 3; ++k) // local->world = T R S (Hittable.cpp:82-102)
	{
		o.l2w[k][0] = scale.x * r[0][k];
		o.l2w[k][1] = scale.y * r[1][k];
		o.l2w[k][2] = scale.z * r[2][k];
		o.l2w[k][3] = p[k];
	}

	// world AABB of the 8 corners of the per-type local extent (Hittable.cpp:139-178)
	float ye[2] = { -1.0f, 1.0f };
	if (d.type == PT_DISK || d.type == PT_QUAD) { ye[0] = -0.01f; ye[1] = 0.01f; }
	else if (d.type == PT_PARABOLOID) { ye[0] = 0.0f; }
	const float xe[2] = { -1.0f, 1.0f }, ze[2] = { -1.0f, 1.0f };
	for (int k = 0; k < 3; ++k) { o.bmin[k] = FLT_MAX; o.bmax[k] = -FLT_MAX; }
	for (int z = 0; z < 2; ++z) for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x)
	{
		const V3 c = { xe[x], ye[y], ze[z] };
		for (int k = 0; k < 3; ++k)
		{
			const float w = dot3(c, V3{ o.l2w[k][0], o.l2w[k][1], o.l2w[k][2] }) + o.l2w[k][3];
			o.bmin[k] = o.bmin[k] < w ? o.bmin[k] : w;
			o.bmax[k] = o.bmax[k] >= w ? o.bmax[k] : w;
		}
	}
}

void computeCamera(const pt_camera_desc &c, CameraDev &out)
{
	auto norm = [](V3 v) { const float inv = 1.0f / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); return V3{ inv * v.x, inv * v.y, inv * v.z }; };
	auto cross = [](V3 u, V3 v) { return V3{ u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x }; };
	const float tanHalf = tanf(c.fovy * 0.5f);
	const V3 origin = { c.position[0], c.position[1], c.position[2] };
	const V3 backward = norm(V3{ origin.x - c.look_at[0], origin.y - c.look_at[1], origin.z - c.look_at[2] });
	const V3 right = norm(cross(V3{ c.up[0], c.up[1], c.up[2] }, backward));
	const V3 up = cross(backward, right);
	const float halfH = tanHalf, halfW = c.aspect * halfH;
	// lowerLeft = -halfW*right + -halfH*up - backward; horizontal = 2*halfW*right; vertical = 2*halfH*up
	const float a = -halfW, b = -halfH, h2 = 2.0f * halfW, v2 = 2.0f * halfH;
	const float r[3] = { right.x, right.y, right.z }, u[3] = { up.x, up.y, up.z }, bw[3] = { backward.x, backward.y, backward.z };
	const float o[3] = { origin.x, origin.y, origin.z };
	for (int k = 0; k < 3; ++k)
	{
		out.origin[k] = o[k];
		out.lowerLeft[k] = (a * r[k] + b * u[k]) - bw[k];
		out.horizontal[k] = h2 * r[k];
		out.vertical[k] = v2 * u[k];
	}
}

namespace
{
struct CamBasis { V3 origin, right, up, backward; };
V3 v3add(V3 a, V3 b) { return V3{ a.x + b.x, a.y + b.y, a.z + b.z }; }
V3 v3scale(V3 a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
V3 v3cross(V3 u, V3 v) { return V3{ u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x }; }
float v3dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 v3norm(V3 v) { const float inv = 1.0f / sqrtf(v3dot(v, v)); return V3{ inv * v.x, inv * v.y, inv * v.z }; }
CamBasis basisOf(const pt_camera_desc &c)
{
	CamBasis b;
	b.origin = V3{ c.position[0], c.position[1], c.position[2] };
	b.backward = v3norm(V3{ b.origin.x - c.look_at[0], b.origin.y - c.look_at[1], b.origin.z - c.look_at[2] }); // Camera.inl:18-20
	b.right = v3norm(v3cross(V3{ c.up[0], c.up[1], c.up[2] }, b.backward));
	b.up = v3cross(b.backward, b.right);
	return b;
}
void writeBack(pt_camera_desc &c, const CamBasis &b)
{
	c.position[0] = b.origin.x; c.position[1] = b.origin.y; c.position[2] = b.origin.z;
	c.look_at[0] = b.origin.x - b.backward.x; c.look_at[1] = b.origin.y - b.backward.y; c.look_at[2] = b.origin.z - b.backward.z;
	c.up[0] = b.up.x; c.up[1] = b.up.y; c.up[2] = b.up.z;
}
// rotateAroundVector, vec3.inl:184-187 (Rodrigues)
V3 rotateAround(V3 v, V3 axis, float cosA, float sinA)
{
	return v3add(v3add(v3scale(v, cosA), v3scale(v3cross(axis, v), sinA)), v3scale(v3scale(axis, v3dot(axis, v)), 1.0f - cosA));
}
} // namespace

void cameraRotate(pt_camera_desc &c, float pitch, float yaw, float roll)
{
	(void)roll; // the reference takes the argument and ignores it (Camera.inl:30-48)
	CamBasis b = basisOf(c);
	const float cosPitch = cosf(-pitch), sinPitch = sinf(-pitch); // around the local x axis
	b.up = rotateAround(b.up, b.right, cosPitch, sinPitch);
	b.backward = rotateAround(b.backward, b.right, cosPitch, sinPitch);
	const float cosYaw = cosf(-yaw), sinYaw = sinf(-yaw); // around the world up axis
	const V3 worldUp = { 0.0f, 1.0f, 0.0f };
	b.right = rotateAround(b.right, worldUp, cosYaw, sinYaw);
	b.up = rotateAround(b.up, worldUp, cosYaw, sinYaw);
	b.backward = rotateAround(b.backward, worldUp, cosYaw, sinYaw);
	writeBack(c, b);
}

void cameraTranslate(pt_camera_desc &c, float x, float y, float z)
{
	CamBasis b = basisOf(c);
	b.origin = v3add(b.origin, v3add(v3add(v3scale(b.right, x), v3scale(b.up, y)), v3scale(b.backward, z))); // Camera.inl:50
	writeBack(c, b);
}

// ---------------------------------------------------------------------------------------------------------------
// BVH build
// ---------------------------------------------------------------------------------------------------------------
namespace
{
constexpr int kBins = 16;
constexpr size_t kWideNode = size_t(1) << 17, kWideChunk = size_t(1) << 14; // nodes over >= 128 k primitives are analysed in 16 k chunks by all cores
constexpr float kTraversalCost = 1.0f; // cost of one two-box node fetch + test, in units of...
constexpr float kHoistFraction = 0.5f; // a primitive whose box has >= this share of the node's box area gets its own leaf at once
constexpr float kPrimCost = 1.5f;      // ...one primitive intersection (quadric tests are dearer than slab tests)

struct Box
{
	float mn[3], mx[3];
	void reset() { for (int k = 0; k < 3; ++k) { mn[k] = FLT_MAX; mx[k] = -FLT_MAX; } }
	void grow(const Box &b) { for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], b.mn[k]); mx[k] = std::max(mx[k], b.mx[k]); } }
	void growPoint(const float *p) { for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); } }
	float area() const
	{
		const float e[3] = { mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2] };
		if (e[0] < 0.0f || e[1] < 0.0f || e[2] < 0.0f) return 0.0f;
		return 2.0f * (e[0] * e[1] + e[0] * e[2] + e[1] * e[2]);
	}
};

struct BuildPrim
{
	Box box;
	float c[3];
	uint32_t index;
};

struct Builder
{
	RecordVector<BuildPrim> &bp;
	RecordVector<Node> &nodes;
	std::vector<uint32_t> subtree; // interior nodes in the subtree of node i, itself included (for the parallel renumbering)
	std::atomic<uint32_t> nextNode{ 0 };
	std::atomic<uint32_t> leafCount{ 0 };
	uint32_t maxLeaf;
	const pt_object_desc *objects;

	Builder(RecordVector<BuildPrim> &b, RecordVector<Node> &n, uint32_t ml, const pt_object_desc *o) : bp(b), nodes(n), subtree(n.size()), maxLeaf(ml), objects(o) {}

	int32_t leafRef(size_t begin, size_t count) const
	{
		const uint32_t t = objects[bp[begin].index].type <= PT_CUBE ? objects[bp[begin].index].type : uint32_t(PT_SPHERE);
		return int32_t(0x80000000u | (t << kLeafTypeShift) | (uint32_t(count) << kLeafCountShift) | uint32_t(begin));
	}

	// returns the child reference for [begin,end), its box in `box`, and the interior depth below it in `depth`
	int32_t build(size_t begin, size_t end, Box &box, uint32_t &depth)
	{
		const size_t n = end - begin;
		// The top of the tree of a large scene: a node over hundreds of thousands of primitives is analysed by ALL cores (chunks of
		// the range as tasks, results merged in chunk order - boxes and counts are order-independent, the "first largest" rule is
		// kept by the merge), not by the one thread that happens to own the node: the first levels were a third of the build.
		const bool wide = n >= kWideNode;
		const size_t nChunks = wide ? std::min<size_t>(64, (n + kWideChunk - 1) / kWideChunk) : 1;
		auto chunkRange = [&](size_t c, size_t &b0, size_t &b1) { b0 = begin + n * c / nChunks; b1 = begin + n * (c + 1) / nChunks; };
		struct Pass1 { Box box, cb, rest, bigBox; float bigArea; size_t big; };
		struct Pass2 { Box bins[3][kBins]; uint32_t counts[3][kBins]; };
		std::vector<Pass1> p1;
		std::vector<Pass2> p2;
		Box cb;
		box.reset();
		cb.reset();
		if (wide)
		{
			p1.resize(nChunks);
			for (size_t c = 0; c < nChunks; ++c)
			{
#pragma omp task shared(p1) firstprivate(c)
				{
					size_t b0, b1;
					chunkRange(c, b0, b1);
					Pass1 &r = p1[c];
					r.box.reset(); r.cb.reset(); r.rest.reset(); r.bigBox.reset();
					r.bigArea = -1.0f; r.big = b0;
					for (size_t i = b0; i < b1; ++i)
					{
						r.box.grow(bp[i].box);
						r.cb.growPoint(bp[i].c);
						const float a = bp[i].box.area();
						if (a > r.bigArea) { if (r.bigArea >= 0.0f) r.rest.grow(r.bigBox); r.bigArea = a; r.big = i; r.bigBox = bp[i].box; }
						else r.rest.grow(bp[i].box);
					}
				}
			}
#pragma omp taskwait
			for (const Pass1 &r : p1) { box.grow(r.box); cb.grow(r.cb); }
		}
		else
			for (size_t i = begin; i < end; ++i) { box.grow(bp[i].box); cb.growPoint(bp[i].c); }
		if (n == 1)
		{
			depth = 0;
			leafCount++;
			return leafRef(begin, n);
		}

		// binned SAH over the centroid bounds
		int bestAxis = -1, bestBin = -1;
		float bestCost = FLT_MAX;
		const float parentArea = std::max(box.area(), 1e-30f);
		if (wide)
		{
			p2.resize(nChunks);
			for (size_t c = 0; c < nChunks; ++c)
			{
#pragma omp task shared(p2, cb) firstprivate(c)
				{
					size_t b0, b1;
					chunkRange(c, b0, b1);
					Pass2 &r = p2[c];
					for (int axis = 0; axis < 3; ++axis)
						for (int b = 0; b < kBins; ++b) { r.bins[axis][b].reset(); r.counts[axis][b] = 0; }
					float k[3];
					for (int axis = 0; axis < 3; ++axis) { const float ext = cb.mx[axis] - cb.mn[axis]; k[axis] = ext > 0.0f ? float(kBins) / ext : 0.0f; }
					for (size_t i = b0; i < b1; ++i)
						for (int axis = 0; axis < 3; ++axis)
						{
							int b = int((bp[i].c[axis] - cb.mn[axis]) * k[axis]);
							b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
							r.counts[axis][b]++;
							r.bins[axis][b].grow(bp[i].box);
						}
				}
			}
#pragma omp taskwait
		}
		for (int axis = 0; axis < 3; ++axis)
		{
			const float ext = cb.mx[axis] - cb.mn[axis];
			if (!(ext > 0.0f)) continue;
			Box bins[kBins];
			uint32_t counts[kBins] = {};
			for (auto &b : bins) b.reset();
			const float k = float(kBins) / ext;
			if (wide)
			{
				for (const Pass2 &r : p2)
					for (int b = 0; b < kBins; ++b) { bins[b].grow(r.bins[axis][b]); counts[b] += r.counts[axis][b]; }
			}
			else
			for (size_t i = begin; i < end; ++i)
			{
				int b = int((bp[i].c[axis] - cb.mn[axis]) * k);
				b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
				counts[b]++;
				bins[b].grow(bp[i].box);
			}
			float rightArea[kBins];
			uint32_t rightCount[kBins];
			Box acc;
			acc.reset();
			uint32_t cnt = 0;
			for (int b = kBins - 1; b > 0; --b) { acc.grow(bins[b]); cnt += counts[b]; rightArea[b] = acc.area(); rightCount[b] = cnt; }
			acc.reset();
			cnt = 0;
			for (int b = 0; b < kBins - 1; ++b)
			{
				acc.grow(bins[b]);
				cnt += counts[b];
				if (cnt == 0 || rightCount[b + 1] == 0) continue;
				const float cost = kTraversalCost + kPrimCost * (float(cnt) * acc.area() + float(rightCount[b + 1]) * rightArea[b + 1]) / parentArea;
				if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
			}
		}

		// one more candidate: the primitive with the largest box on its own against all the others.  Centroid binning can
		// never isolate a primitive that spans the node (a floor under hundreds of small objects): it drags its huge box
		// down every level of the subtree that happens to hold its centroid.
		size_t isolate = end;
		if (n >= 3)
		{
			size_t big = begin;
			float bigArea = -1.0f;
			Box rest;
			rest.reset();
			if (wide)
			{
				for (const Pass1 &r : p1) if (r.bigArea > bigArea) { bigArea = r.bigArea; big = r.big; } // strict: the first largest, as below
				for (const Pass1 &r : p1) { rest.grow(r.rest); if (r.big != big) rest.grow(r.bigBox); }
			}
			else
			{
				for (size_t i = begin; i < end; ++i) { const float a = bp[i].box.area(); if (a > bigArea) { bigArea = a; big = i; } }
				for (size_t i = begin; i < end; ++i) if (i != big) rest.grow(bp[i].box);
			}
			const float cost = kTraversalCost + kPrimCost * (bigArea + float(n - 1) * rest.area()) / parentArea;
			// the greedy SAH cost (children costed as leaves) cannot see that damage, so a primitive covering most of the
			// node is hoisted outright
			if (cost < bestCost || bigArea >= kHoistFraction * parentArea) { bestCost = std::min(bestCost, cost); isolate = big; }
		}

		if (n <= maxLeaf && ((bestAxis < 0 && isolate == end) || bestCost >= kPrimCost * float(n)))
		{
			// the primitives of a leaf in the order quadrics, flat shapes, cubes (the three branches of intersectLocal): the lanes
			// of a warp that test their leaves' i-th primitives together then take the same branch more often.  (Ties are broken by
			// scene index in the kernels, so the order inside a leaf cannot change a result.)
			auto shapeClass = [this](const BuildPrim &p) -> int
			{
				const uint32_t t = objects[p.index].type;
				return (t == PT_DISK || t == PT_QUAD) ? 1 : (t == PT_CUBE ? 2 : 0);
			};
			std::sort(bp.begin() + begin, bp.begin() + end, [&](const BuildPrim &a, const BuildPrim &b)
				{ const int ca = shapeClass(a), cb2 = shapeClass(b); return ca < cb2 || (ca == cb2 && a.index < b.index); });
			depth = 0;
			leafCount++;
			return leafRef(begin, n);
		}

		size_t mid = begin;
		if (isolate != end)
		{
			std::swap(bp[begin], bp[isolate]);
			mid = begin + 1;
		}
		else if (bestAxis >= 0)
		{
			const float ext = cb.mx[bestAxis] - cb.mn[bestAxis];
			const float k = float(kBins) / ext, lo = cb.mn[bestAxis];
			const int axis = bestAxis, bin = bestBin;
			auto it = std::partition(bp.begin() + begin, bp.begin() + end, [=](const BuildPrim &p)
				{
					int b = int((p.c[axis] - lo) * k);
					b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
					return b <= bin;
				});
			mid = size_t(it - bp.begin());
		}
		if (mid == begin || mid == end)
		{
			// all centroids coincide (or binning failed): split the range in the middle by index along the widest axis
			int axis = 0;
			const float e[3] = { box.mx[0] - box.mn[0], box.mx[1] - box.mn[1], box.mx[2] - box.mn[2] };
			if (e[1] > e[axis]) axis = 1;
			if (e[2] > e[axis]) axis = 2;
			mid = (begin + end) / 2;
			std::nth_element(bp.begin() + begin, bp.begin() + mid, bp.begin() + end, [axis](const BuildPrim &a, const BuildPrim &b)
				{ return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.index < b.index); });
		}

		const uint32_t nodeIndex = nextNode++;
		Box lb, rb;
		uint32_t ld = 0, rd = 0;
		int32_t lc, rc;
		if (n > 8192)
		{
#pragma omp task shared(lb, ld, lc) default(shared)
			lc = build(begin, mid, lb, ld);
			rc = build(mid, end, rb, rd);
#pragma omp taskwait
		}
		else
		{
			lc = build(begin, mid, lb, ld);
			rc = build(mid, end, rb, rd);
		}
		Node &nd = nodes[nodeIndex];
		for (int k = 0; k < 3; ++k) { nd.f[k] = lb.mn[k]; nd.f[3 + k] = lb.mx[k]; nd.f[6 + k] = rb.mn[k]; nd.f[9 + k] = rb.mx[k]; }
		nd.child[0] = lc;
		nd.child[1] = rc;
		nd.pad[0] = nd.pad[1] = 0;
		subtree[nodeIndex] = 1u + (lc >= 0 ? subtree[size_t(lc)] : 0u) + (rc >= 0 ? subtree[size_t(rc)] : 0u);
		depth = 1 + std::max(ld, rd);
		return int32_t(nodeIndex);
	}
};
// ---------------------------------------------------------------------------------------------------------------
// LBVH: the fast builder for very large scenes ("LBVH / PLOC"; option "bvh_builder").  Morton order of the
// centroids (63-bit keys, 21 bits per axis), the radix tree of Karras 2012 over the sorted keys - every internal node finds its
// range and split on its own, so all cores (or all threads of a GPU: the code is data-parallel throughout) work from the first
// instruction - and one bottom-up pass that fits the boxes and records depth and subtree sizes.  One primitive per leaf.  The
// tree has the builder's usual output form (Builder::nodes with child boxes as min / max, `subtree`), so renumbering, box
// conversion and record packing are shared with the SAH builder.  Against SAH: the build is several times faster, the
// traversal a little slower (measured: DESIGN.md section 7).
// ---------------------------------------------------------------------------------------------------------------
inline uint64_t spread21(uint64_t v) // 21 bits -> every third bit
{
	v &= 0x1fffffull;
	v = (v | v << 32) & 0x1f00000000ffffull;
	v = (v | v << 16) & 0x1f0000ff0000ffull;
	v = (v | v << 8) & 0x100f00f00f00f00full;
	v = (v | v << 4) & 0x10c30c30c30c30c3ull;
	v = (v | v << 2) & 0x1249249249249249ull;
	return v;
}
struct LbvhKey { uint64_t key; uint32_t index; };

// returns the root reference (internal node 0, or a leaf reference when there is a single primitive)
int32_t buildLbvh(Builder &b, size_t begin, size_t end, Box &rootBox, uint32_t &depth)
{
	RecordVector<BuildPrim> &bp = b.bp;
	const size_t m = end - begin;
	rootBox.reset();
	if (m == 1)
	{
		rootBox = bp[begin].box;
		depth = 0;
		b.leafCount++;
		return b.leafRef(begin, 1);
	}
	// centroid bounds
	Box cb;
	cb.reset();
#pragma omp parallel
	{
		Box mine;
		mine.reset();
#pragma omp for schedule(static) nowait
		for (long i = long(begin); i < long(end); ++i) mine.growPoint(bp[size_t(i)].c);
#pragma omp critical(ptb_lbvh_bounds)
		cb.grow(mine);
	}
	// Morton keys, sorted (ties by index: the order - and so the tree - is the same whatever the thread count)
	RecordVector<LbvhKey> keys(m);
	{
		// ONE scale for the three axes (the widest extent fills the 21 bits): the cells of the Morton grid are cubes, so a flat scene -
		// a million objects on a plane - is split along the plane's two axes only (its third coordinate has no bits that differ, and
		// the radix tree skips bits that do not differ).  With one scale per axis a third of the splits went along the thin axis: boxes
		// overlapping everywhere, the 1 M-object scene rendered 5.6x slower than with the SAH tree.
		float widest = 0.0f;
		for (int k = 0; k < 3; ++k) widest = std::max(widest, cb.mx[k] - cb.mn[k]);
		float scale[3];
		for (int k = 0; k < 3; ++k) scale[k] = widest > 0.0f ? 2097151.0f / widest : 0.0f;
#pragma omp parallel for schedule(static)
		for (long i = 0; i < long(m); ++i)
		{
			const BuildPrim &p = bp[begin + size_t(i)];
			uint64_t q[3];
			for (int k = 0; k < 3; ++k)
			{
				const float f = (p.c[k] - cb.mn[k]) * scale[k];
				q[k] = uint64_t(f < 0.0f ? 0.0f : (f > 2097151.0f ? 2097151.0f : f));
			}
			keys[size_t(i)].key = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
			keys[size_t(i)].index = uint32_t(i);
		}
	}
	auto keyLess = [](const LbvhKey &x, const LbvhKey &y) { return x.key < y.key || (x.key == y.key && x.index < y.index); };
#ifdef _OPENMP
	__gnu_parallel::sort(keys.begin(), keys.end(), keyLess);
#else
	std::sort(keys.begin(), keys.end(), keyLess);
#endif
	// the primitives in Morton order
	{
		RecordVector<BuildPrim> sorted(m);
#pragma omp parallel for schedule(static)
		for (long i = 0; i < long(m); ++i) sorted[size_t(i)] = bp[begin + keys[size_t(i)].index];
#pragma omp parallel for schedule(static)
		for (long i = 0; i < long(m); ++i) bp[begin + size_t(i)] = sorted[size_t(i)];
	}
	// common prefix of keys i and j (position in the sorted order breaks ties between equal keys), -1 outside the range
	const long n = long(m);
	auto delta = [&keys, n](long i, long j) -> int
	{
		if (j < 0 || j >= n) return -1;
		const uint64_t x = keys[size_t(i)].key ^ keys[size_t(j)].key;
		if (x != 0) return __builtin_clzll(x);
		return 64 + __builtin_clz(uint32_t(i) ^ uint32_t(j));
	};
	// Karras 2012, figure 4: internal node i covers [first, last], split behind `split`; children: internal node `split` / `split + 1`
	// unless the half is a single key (then the leaf).  parent pointers: internal nodes 0 .. n - 2, leaves n - 1 + position
	const size_t internal = m - 1;
	b.nextNode = uint32_t(internal);
	std::vector<int32_t> parent(internal + m, -1);
	std::vector<int32_t> kids(2 * internal);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < long(internal); ++i)
	{
		const int d = delta(i, i + 1) - delta(i, i - 1) > 0 ? 1 : -1;
		const int dMin = delta(i, i - d);
		long lMax = 2;
		while (delta(i, i + lMax * d) > dMin) lMax *= 2;
		long l = 0;
		for (long t = lMax / 2; t >= 1; t /= 2)
			if (delta(i, i + (l + t) * d) > dMin) l += t;
		const long j = i + l * d;
		const int dNode = delta(i, j);
		long sOff = 0;
		for (long t = (l + 1) / 2;; t = (t + 1) / 2)
		{
			if (delta(i, i + (sOff + t) * d) > dNode) sOff += t;
			if (t == 1) break;
		}
		const long split = i + sOff * d + std::min(d, 0);
		const long first = std::min(i, j), last = std::max(i, j);
		const int32_t left = first == split ? int32_t(internal + size_t(split)) : int32_t(split);           // (>= internal: a leaf, by position)
		const int32_t right = last == split + 1 ? int32_t(internal + size_t(split + 1)) : int32_t(split + 1);
		kids[2 * size_t(i)] = left;
		kids[2 * size_t(i) + 1] = right;
		parent[size_t(left)] = int32_t(i);
		parent[size_t(right)] = int32_t(i);
	}
	// bottom-up: the second thread to arrive at a node finds both children complete, fits the node and goes on
	struct Fit { Box box; uint32_t depth, subtree; };
	RecordVector<Fit> fit(internal);
	std::vector<std::atomic<uint32_t>> arrived(internal);
#pragma omp parallel for schedule(static)
	for (long i = 0; i < long(internal); ++i) arrived[size_t(i)].store(0, std::memory_order_relaxed);
	std::atomic<uint32_t> rootDepth{ 0 };
#pragma omp parallel for schedule(static)
	for (long leaf = 0; leaf < n; ++leaf)
	{
		int32_t node = parent[internal + size_t(leaf)];
		while (node >= 0)
		{
			if (arrived[size_t(node)].fetch_add(1, std::memory_order_acq_rel) == 0) break; // the first to arrive leaves the rest to the second
			Node &nd = b.nodes[size_t(node)];
			Fit &f = fit[size_t(node)];
			f.box.reset();
			f.depth = 0;
			f.subtree = 1;
			for (int c = 0; c < 2; ++c)
			{
				const int32_t k = kids[2 * size_t(node) + size_t(c)];
				Box cbx;
				if (k >= int32_t(internal))
				{
					const size_t pos = begin + size_t(k) - internal;
					cbx = bp[pos].box;
					nd.child[c] = b.leafRef(pos, 1);
				}
				else
				{
					cbx = fit[size_t(k)].box;
					f.depth = std::max(f.depth, fit[size_t(k)].depth);
					f.subtree += fit[size_t(k)].subtree;
					nd.child[c] = k;
				}
				for (int a = 0; a < 3; ++a) { nd.f[6 * c + a] = cbx.mn[a]; nd.f[6 * c + 3 + a] = cbx.mx[a]; }
				f.box.grow(cbx);
			}
			f.depth += 1;
			nd.pad[0] = nd.pad[1] = 0;
			b.subtree[size_t(node)] = f.subtree;
			if (node == 0) rootDepth.store(f.depth);
			node = parent[size_t(node)];
		}
	}
	b.leafCount += uint32_t(m);
	rootBox = fit[0].box;
	depth = rootDepth.load();
	return 0;
}
} // namespace

bool compileScene(size_t count, const pt_object_desc *objects, uint32_t maxLeaf, CompiledScene &out, std::string &err, uint32_t maxGlobal, const ObjectXform *given,
                  int builder)
{
	out = CompiledScene();
	static const bool ptbTiming = getenv("PTB_TIMING") != nullptr;
	auto tNow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double tLast = tNow();
	auto lap = [&](const char *what) { if (ptbTiming) { const double t = tNow(); fprintf(stderr, "  compileScene %-14s %7.1f ms\n", what, (t - tLast) * 1e3); tLast = t; } };
	if (count == 0) { err = "empty scene"; return false; }
	if (count > kLeafStartMask) { err = "too many objects"; return false; }
	maxLeaf = std::max(1u, std::min(maxLeaf, kMaxLeafPrims));

	RecordVector<ObjectXform> xf(count); // (both filled by the parallel loop below)
	RecordVector<BuildPrim> bp(count);
#pragma omp parallel for schedule(static) if (count > 4096)
	for (long i = 0; i < long(count); ++i)
	{
		if (given) xf[i] = given[i];
		else computeObjectXform(objects[i], xf[i]);
		for (int k = 0; k < 3; ++k)
		{
			bp[i].box.mn[k] = xf[i].bmin[k];
			bp[i].box.mx[k] = xf[i].bmax[k];
			bp[i].c[k] = 0.5f * (xf[i].bmin[k] + xf[i].bmax[k]);
		}
		bp[i].index = uint32_t(i);
	}

	lap("xforms");
	// hoist the primitives that span the scene (pt_types.h kMaxGlobalPrims): largest first, at the front of the order
	Box sceneBox;
	sceneBox.reset();
	for (size_t i = 0; i < count; ++i) sceneBox.grow(bp[i].box);
	size_t nGlobal = 0;
	{
		const float need = kGlobalAreaFraction * sceneBox.area();
		const size_t cap = std::min<size_t>(std::min<uint32_t>(maxGlobal, kMaxGlobalPrims), count - 1); // keep one primitive for the tree
		while (nGlobal < cap)
		{
			size_t big = nGlobal;
			float bigArea = -1.0f;
			for (size_t i = nGlobal; i < count; ++i) { const float a = bp[i].box.area(); if (a > bigArea) { bigArea = a; big = i; } }
			if (!(bigArea >= need) || !(need > 0.0f)) break;
			std::swap(bp[nGlobal], bp[big]);
			++nGlobal;
		}
	}
	out.globalCount = uint32_t(nGlobal);

	lap("hoist");
	out.nodes.resize(std::max<size_t>(count, 2) - 1 + 1); // (not cleared: the builder writes every node it hands out)
	lap("alloc nodes");
	Builder b(bp, out.nodes, maxLeaf, objects);
	Box rootBox;
	uint32_t depth = 0;
	int32_t root;
	// the builder: SAH unless asked otherwise; kBuilderAuto picks the LBVH for very large scenes, where the build is most of the load
	bool lbvh = builder == kBuilderLbvh || (builder == kBuilderAuto && count >= kLbvhAutoCount);
	if (lbvh)
	{
		root = buildLbvh(b, nGlobal, count, rootBox, depth);
		// a radix tree is as deep as the keys make it (up to 63 + 32 levels for centroids that crowd towards one point): one that the
		// traversal stack cannot hold is thrown away and the SAH builder, whose depth is bounded by its median fall-back, takes over
		if (depth + 2 > uint32_t(kStackSize))
		{
			lbvh = false;
			b.nextNode = 0;
			b.leafCount = 0;
		}
	}
	if (!lbvh)
	{
#pragma omp parallel if (count > 8192)
#pragma omp single
		root = b.build(nGlobal, count, rootBox, depth);
	}

	lap("build");
	if (root < 0)
	{
		// the whole scene is one leaf: give it a root node whose second child is empty
		Node &nd = out.nodes[0];
		// the empty child's box is a point at +FLT_MAX: its slab interval is degenerate (near == far) for every ray, so the
		// strict near < far test never passes.  (An inverted box would NOT work: the min/max slab test un-inverts it.)
		for (int k = 0; k < 3; ++k) { nd.f[k] = rootBox.mn[k]; nd.f[3 + k] = rootBox.mx[k]; nd.f[6 + k] = FLT_MAX; nd.f[9 + k] = FLT_MAX; }
		// (converted below to centre +FLT_MAX / half extent -1)
		nd.child[0] = root;
		nd.child[1] = kEmptyChild;
		nd.pad[0] = nd.pad[1] = 0;
		b.nextNode = 1;
		depth = 1;
	}
	else if (root != 0)
	{
		err = "internal: root is not node 0";
		return false;
	}
	out.nodes.resize(b.nextNode.load());
	if (depth + 2 > uint32_t(kStackSize)) { err = "BVH deeper than the traversal stack"; return false; }
	// min/max -> centre / half extent, padded outwards: the traversal computes t_c = c*inv - o*inv in fp32, so the box
	// must absorb a few ulp of |c|, of its own size and of the scene scale (ray origins) to stay conservative.  Applied to every
	// node on its way into its final place (the renumbering copy below; the records appended behind the tree).
	float boxScale = 0.0f;
	for (int k = 0; k < 3; ++k) boxScale = std::max(boxScale, std::max(fabsf(sceneBox.mn[k]), fabsf(sceneBox.mx[k])));
	auto toCentreHalf = [boxScale](Node &nd)
	{
		float mm[12]; // as the builder left them: child c = min[3] max[3] at 6 * c
		memcpy(mm, nd.f, sizeof mm);
		for (int c = 0; c < 2; ++c)
		{
			const float *f = mm + 6 * c;
			for (int k = 0; k < 3; ++k)
			{
				if (nd.child[c] == kEmptyChild)
				{
					nd.f[nodeF(c, 0, k)] = FLT_MAX;
					nd.f[nodeF(c, 1, k)] = -1.0f;
					continue;
				}
				const double mn = f[k], mx = f[3 + k];
				const float ctr = float(0.5 * (mn + mx));
				double h = std::max(mx - double(ctr), double(ctr) - mn); // covers the rounding of the centre
				h += 4.0e-7 * (fabs(double(ctr)) + h + double(boxScale)) + 1e-30;
				nd.f[nodeF(c, 0, k)] = ctr;                              // interleaved layout of pt_types.h
				nd.f[nodeF(c, 1, k)] = nextafterf(float(h), FLT_MAX);
			}
		}
	};
	// Renumber the nodes: the first kTopOrderNodes in BREADTH-FIRST order (the levels every ray walks sit in a few
	// consecutive cache lines), the rest depth-first (subtrees contiguous).  Also makes the layout independent of the
	// order in which the parallel build tasks ran.  (Staging that breadth-first prefix in shared memory for scenes that do
	// not fit was measured and dropped: L1 already holds it, and the generic-address loads cost more - 7.2 vs 7.8 Grays/s.)
	if (out.nodes.size() > 1)
	{
		const size_t n = out.nodes.size();
		std::vector<int32_t> newIndex(n, -1);
		std::vector<int32_t> queue;
		queue.push_back(0);
		size_t head = 0, numbered = 0;
		while (head < queue.size() && numbered < kTopOrderNodes)
		{
			const int32_t i = queue[head++];
			newIndex[i] = int32_t(numbered++);
			for (int c = 0; c < 2; ++c)
				if (out.nodes[i].child[c] >= 0) queue.push_back(out.nodes[i].child[c]);
		}
		// the frontier (queued but not numbered) and everything below it: depth-first, in queue order.  Every frontier subtree
		// takes a block of indices whose start follows from the subtree sizes the builder recorded, so the subtrees are numbered
		// by all cores at once
		const size_t nFront = queue.size() - head;
		std::vector<size_t> startOf(nFront + 1);
		startOf[0] = numbered;
		for (size_t f = 0; f < nFront; ++f) startOf[f + 1] = startOf[f] + b.subtree[size_t(queue[head + f])];
		if (startOf[nFront] != n) { err = "internal: BVH renumbering lost nodes"; return false; }
#pragma omp parallel for schedule(dynamic, 16) if (n > 8192)
		for (long f = 0; f < long(nFront); ++f)
		{
			int32_t stack[2 * kStackSize + 8];
			int sp = 0;
			stack[sp++] = queue[head + size_t(f)];
			size_t next = startOf[size_t(f)];
			while (sp > 0)
			{
				const int32_t i = stack[--sp];
				newIndex[i] = int32_t(next++);
				for (int c = 1; c >= 0; --c)
					if (out.nodes[i].child[c] >= 0 && sp < int(sizeof stack / sizeof stack[0])) stack[sp++] = out.nodes[i].child[c];
			}
		}
		RecordVector<Node> renum(n);
#pragma omp parallel for schedule(static) if (n > 8192)
		for (long i = 0; i < long(n); ++i)
		{
			Node nd = out.nodes[size_t(i)];
			for (int c = 0; c < 2; ++c)
				if (nd.child[c] >= 0) nd.child[c] = newIndex[nd.child[c]];
			toCentreHalf(nd);
			renum[size_t(newIndex[size_t(i)])] = nd;
		}
		out.nodes.swap(renum);
	}
	else toCentreHalf(out.nodes[0]);
	lap("renumber");
	// the hoisted primitives' boxes, two per record, behind the tree: the pixel-beam walk decides with them which hoisted
	// primitives the camera rays of a pixel have to test at all
	out.treeNodeCount = uint32_t(out.nodes.size());
	for (size_t g = 0; g < nGlobal; g += 2)
	{
		Node nd;
		for (int c = 0; c < 2; ++c)
		{
			float *f = nd.f + 6 * c;
			if (g + c < nGlobal)
			{
				for (int k = 0; k < 3; ++k) { f[k] = bp[g + c].box.mn[k]; f[3 + k] = bp[g + c].box.mx[k]; }
				nd.child[c] = b.leafRef(g + c, 1);
			}
			else
			{
				for (int k = 0; k < 3; ++k) { f[k] = FLT_MAX; f[3 + k] = FLT_MAX; }
				nd.child[c] = kEmptyChild;
			}
		}
		nd.pad[0] = nd.pad[1] = 0;
		toCentreHalf(nd);
		out.nodes.push_back(nd);
	}
	out.depth = depth;
	out.leafCount = b.leafCount.load();
	for (int k = 0; k < 3; ++k) { out.sceneMin[k] = sceneBox.mn[k]; out.sceneMax[k] = sceneBox.mx[k]; }

	out.prims.resize(count);
	out.mats.resize(count);
#pragma omp parallel for schedule(static) if (count > 8192)
	for (long i = 0; i < long(count); ++i)
	{
		const uint32_t src = bp[i].index;
		const pt_object_desc &d = objects[src];
		Prim &p = out.prims[i];
		memcpy(p.row0, xf[src].w2l[0], 16);
		memcpy(p.row1, xf[src].w2l[1], 16);
		memcpy(p.row2, xf[src].w2l[2], 16);
		const uint32_t type = d.type <= PT_CUBE ? d.type : uint32_t(PT_SPHERE);
		// quadric coefficients: the template arguments of Hittable.inl:151 sphere <1,1,1,..,-1>, :176 cylinder <1,0,1,..,-1>, :242 cone <1,-1,1>,
		// :273 paraboloid <1,0,1,..,H=-1>
		p.qB = type == PT_SPHERE ? 1.0f : (type == PT_CONE ? -1.0f : 0.0f);
		p.qH = type == PT_PARABOLOID ? -1.0f : 0.0f;
		p.qJ = (type == PT_SPHERE || type == PT_CYLINDER) ? -1.0f : 0.0f;
		p.packed = src | (type << kPrimTypeShift) | (d.material.texture != 0 ? kPrimTextured : 0u) | ((type == PT_DISK || type == PT_QUAD) ? kPrimFlat : 0u) |
		           (type == PT_CUBE ? kPrimCube : 0u) | (type == PT_DISK ? kPrimDisk : 0u) | (type == PT_SPHERE ? kPrimSphere : 0u);
		Mat &m = out.mats[i];
		memcpy(m.baseColor, d.material.base_color, 12);
		memcpy(m.emissive, d.material.emissive, 12);
		m.roughness = d.material.roughness < 0.04f ? 0.04f : d.material.roughness; // Material.inl:12
		m.metalness = d.material.metalness;
		m.texture = d.material.texture;
		m.type = d.material.type <= PT_LAMBERT_GGX ? d.material.type : uint32_t(PT_LAMBERT);
		m.pad[0] = m.pad[1] = 0;
	}
	lap("records");
	return true;
}

} // namespace ptb
