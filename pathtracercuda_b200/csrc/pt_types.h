// pt_types.h — device-side data layout shared by the host scene compiler and the CUDA kernels.
// All records are multiples of 16 bytes so every access is a 128-bit load (LDS.128 from the shared-memory copy,
// LDG.E.128.CONSTANT from HBM/L2 for scenes that do not fit).
#pragma once
#include <cstdint>

namespace ptb
{

// One BVH node = the boxes of BOTH children + their references (64 B, one 128-bit x4 fetch tests two boxes).
// Replaces the reference's 32 B single-box node (BVH.h:6-11), whose traversal needs one dependent fetch per box.
//   Boxes are stored as CENTRE / HALF EXTENT (padded outwards by a few ulp of the scene scale): per axis
//   t_c = c*inv - o*inv, near = t_c - h*|inv|, far = t_c + h*|inv| - FMAs instead of min/max pairs - and the two children
//   are INTERLEAVED so that every 64-bit register pair of the 128-bit loads holds the same quantity of child 0 and child 1:
//   sm_100's packed FFMA2 (fma.rn.f32x2) then does both children in ONE instruction (9 FFMA2 per node instead of 18 FFMA):
//     f[0..5]  = (c0.x c1.x) (c0.y c1.y) (c0.z c1.z)      centres
//     f[6..11] = (h0.x h1.x) (h0.y h1.y) (h0.z h1.z)      half extents          -> nodeF(child, half?, axis)
//   child[k] >= 0 : index of an interior node
//   child[k] <  0 : leaf; bits 0..23 = first primitive (BVH order), bits 24..27 = primitive count (1..15),
//                   bits 28..30 = shape type of the first primitive
//   an EMPTY child has child = kEmptyChild, centre +FLT_MAX and half extent -1 (near > far for every ray: never hit)
struct alignas(16) Node
{
	float f[12];
	int32_t child[2];
	uint32_t pad[2];
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");
// index into Node::f of (child 0/1, centre (0) or half extent (1), axis 0..2)
constexpr int nodeF(int child, int half, int axis) { return 6 * half + 2 * axis + child; }

constexpr int32_t kLeafBit = int32_t(0x80000000u);
constexpr int kLeafCountShift = 24;
constexpr int kLeafTypeShift = 28;
constexpr uint32_t kLeafStartMask = (1u << kLeafCountShift) - 1u;
constexpr int32_t kEmptyChild = kLeafBit; // leaf with count 0
constexpr uint32_t kMaxLeafPrims = 15;
constexpr int kStackSize = 48;
// Primitives whose box covers a large share of the scene box (a floor, the walls of a room) are kept OUT of the BVH and
// tested first by every ray: the whole warp runs the same test on the same record (no divergence, broadcast loads), the
// traversal starts with a tight t_max, and the tree is not polluted by boxes that overlap everything.
constexpr uint32_t kMaxGlobalPrims = 8;
constexpr size_t kTopOrderNodes = 4096; // nodes[0..4096) are numbered breadth-first, the rest depth-first (scene_compile.cpp)
constexpr float kGlobalAreaFraction = 0.25f;

// One primitive = world->local 3x4 (the reference's Hittable rows, Hittable.h:22-24) + one quad of shape data (64 B).
// The 40 B material the reference embeds in every 96 B Hittable (Hittable.h:25) lives in its own table: it is
// only needed once per path segment (at the closest hit), not once per candidate.
// The fourth quad: the quadric's coefficients as the floats the intersection routine multiplies with (B, H, J of
// A x^2 + B y^2 + C z^2 + H y + J = 0 with A = C = 1; zero for the other shapes) and ONE word with everything discrete, laid out
// so that every decision of the primitive test is a single-bit test: no compare chains on the type, no int -> float conversions.
struct alignas(16) Prim
{
	float row0[4];
	float row1[4];
	float row2[4];
	float qB, qH, qJ;
	uint32_t packed; // kPrim* below
};
constexpr uint32_t kPrimSceneMask = 0x00ffffffu; // bits 0..23: index in the caller's object array (tie-break + primary-pass output)
constexpr int kPrimTypeShift = 24;               // bits 24..26: PT_SPHERE .. PT_CUBE
constexpr uint32_t kPrimTextured = 1u << 27;     // needs u, v
constexpr uint32_t kPrimFlat = 1u << 28;         // disk or quad
constexpr uint32_t kPrimCube = 1u << 29;
constexpr uint32_t kPrimDisk = 1u << 30;
constexpr uint32_t kPrimSphere = 1u << 31;
constexpr uint32_t kMaxSceneObjects = kPrimSceneMask + 1u; // (a leaf reference holds 24 bits of primitive index as well)

static_assert(sizeof(Prim) == 64, "Prim must be 64 bytes");

// Material (reference Material.h:22-27, 40 B) padded to 48 B, with roughness already clamped (Material.inl:12).
struct alignas(16) Mat
{
	float baseColor[3];
	float roughness;
	float emissive[3];
	float metalness;
	uint32_t texture; // 1-based handle, 0 = none
	uint32_t type;    // PT_LAMBERT / PT_GGX / PT_LAMBERT_GGX
	uint32_t pad[2];
};
static_assert(sizeof(Mat) == 48, "Mat must be 48 bytes");

// Packed texture descriptor for the read-only path: LDR = RGBA8 (one 32-bit texel), HDR = float4 (one 128-bit texel)
struct TexDesc
{
	const void *texels;
	uint32_t width;
	uint32_t height;
	uint32_t isHdr;
	uint32_t pad;
	// the same texels as a CUDA texture object (array + linear filter, wrap U / clamp V, normalised coordinates: the reference's
	// own setup, Pathtracer.cpp:259-288): ONE TEX instruction per tap instead of ~120 for four loads + fp32 filtering.
	// 0 = filter in software from `texels` (option tex_unit=0, and the CPU emulation).
	unsigned long long texObj;
};
static_assert(sizeof(TexDesc) == 32, "TexDesc must be 32 bytes");

struct CameraDev
{
	float origin[3];
	float lowerLeft[3];
	float horizontal[3];
	float vertical[3];
};

} // namespace ptb
