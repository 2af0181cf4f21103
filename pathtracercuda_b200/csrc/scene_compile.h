// scene_compile.h — host scene compiler: object descriptions -> flat device arrays (nodes, primitives, materials).
#pragma once
#include "../../include/pt_b200.h"
#include "pt_types.h"
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace ptb
{

// std::vector without the zero fill: resize() of these record arrays (tens to hundreds of megabytes for a million objects) default-
// initialises - every element is written by the compiler before it is read - instead of clearing memory on one thread first
template <typename T>
struct DefaultInitAllocator : std::allocator<T>
{
	template <typename U> struct rebind { using other = DefaultInitAllocator<U>; };
	DefaultInitAllocator() = default;
	template <typename U> DefaultInitAllocator(const DefaultInitAllocator<U> &) {}
	template <typename U> void construct(U *p) { ::new (static_cast<void *>(p)) U; }
	template <typename U, typename... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <typename T> using RecordVector = std::vector<T, DefaultInitAllocator<T>>;

struct CompiledScene
{
	RecordVector<Node> nodes;  // nodes[0..treeNodeCount): the tree; then one record per PAIR of hoisted primitives (their boxes + leaf
	                          // references), which no tree node points to: only the pixel-beam walk reads them (trace_device.cuh)
	uint32_t treeNodeCount = 0;
	RecordVector<Prim> prims; // BVH order
	RecordVector<Mat> mats;   // BVH order (parallel to prims)
	uint32_t depth = 0;      // interior-node depth (max traversal stack = depth)
	uint32_t leafCount = 0;
	uint32_t globalCount = 0; // prims[0..globalCount) are tested by every ray before the traversal and are not in the BVH
	float sceneMin[3] = { 0, 0, 0 };
	float sceneMax[3] = { 0, 0, 0 };
};

struct ObjectXform
{
	float w2l[3][4]; // world -> local rows
	float l2w[3][4]; // local -> world rows
	float bmin[3], bmax[3];
};

// CpuHittable ctor equivalent (reference Hittable.cpp:115-179): rows + world AABB, float arithmetic in the
// reference's operation order so the rows are bit-identical to the reference's.
void computeObjectXform(const pt_object_desc &d, ObjectXform &out);

// Build the BVH (binned SAH, kBins bins, centroid bounds, leaf when cheaper) and flatten it.
// maxLeaf in [1, kMaxLeafPrims].  Returns false and sets err on failure (depth over kStackSize).
// `given` != nullptr: the objects' transforms and world boxes as the caller has them (pt_set_scene_xform: only w2l, bmin, bmax
// are read) instead of deriving them from position / rotation / scale.
// `builder`: kBuilderSah (binned SAH, the default), kBuilderLbvh (Morton order + Karras' radix tree: a several times faster build, a
// somewhat slower traversal), kBuilderAuto (LBVH from kLbvhAutoCount objects).
enum { kBuilderSah = 0, kBuilderLbvh = 1, kBuilderAuto = 2 };
constexpr size_t kLbvhAutoCount = size_t(1) << 19;
bool compileScene(size_t count, const pt_object_desc *objects, uint32_t maxLeaf, CompiledScene &out, std::string &err, uint32_t maxGlobal = kMaxGlobalPrims,
                  const ObjectXform *given = nullptr, int builder = kBuilderSah);

// Camera ctor + update() equivalent (reference Camera.inl:4-23,54-62)
void computeCamera(const pt_camera_desc &c, CameraDev &out);
// Camera::rotate / Camera::translate (reference Camera.inl:30-52) on the POD camera: the basis (right, up, backward) is
// rebuilt from the description as the ctor does, moved as the reference moves it, and written back as
// look_at = position - backward, up = the new up vector (the ctor then reproduces the same basis).
void cameraRotate(pt_camera_desc &c, float pitch, float yaw, float roll);
void cameraTranslate(pt_camera_desc &c, float x, float y, float z);

} // namespace ptb
