// Pathtracer_b200.cpp — the reference-side binding: a body for the reference's own `class Pathtracer`
// (PathtracerCUDA/src/pathtracer/Pathtracer.h:12-69, included UNMODIFIED) that forwards every public method to the C ABI of
// libpt_b200.so (include/pt_b200.h).  A maintainer of the reference builds main.cpp, SceneLoader.cpp, Hittable.cpp and the
// window / image-writer sources exactly as before and links THIS file + -lpt_b200 instead of Pathtracer.cpp, BVH.cpp and
// kernels/*.cu.  oracle/Makefile target _ref/ref_main_b200 does precisely that, and tests/test_gpu_parity.py renders both
// bundled scenes through it: the reference's own front end (argument parsing, JSON loader, 8-spp render loop, PNG / HDR
// writers) on our back end.
//
// What the shim has to work around, all of it inside this file:
//   * the class has no member to spare: the pt_context pointer lives in m_gpuAccumBuffer (a float4* the reference uses for its
//     own device buffer; nothing outside Pathtracer.cpp touches it);
//   * a CpuHittable keeps no position / rotation / scale, only the world->local rows, the AABB, the type and the material, and
//     hands them out solely as a Hittable (getGpuHittable(), Hittable.h:22-27: rows at byte 0, Material at 48, type at 88;
//     Material.h:22-27: baseColor 0, roughness 12, emissive 16, metalness 28, textureIndex 32, type 36) - read by offset, with
//     static_asserts on the sizes, and passed on through pt_set_scene_xform;
//   * the Camera is passed as its four ray vectors (public members, Camera.h:15-18) through pt_render_vectors: no re-derivation.
#include "pathtracer/Pathtracer.h"
#include "pathtracer/Camera.h"
#include "pathtracer/Hittable.h"
#include "pt_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static_assert(sizeof(Hittable) == 96, "Hittable layout changed: rows @0, Material @48, type @88 (Hittable.h:22-27)");
static_assert(sizeof(Material) == 40, "Material layout changed (Material.h:22-27)");

static pt_context *ctxOf(float4 *slot) { return reinterpret_cast<pt_context *>(slot); }
static void check(int rc, const char *what)
{
	// the reference prints the CUDA error and exits with EXIT_FAILURE (Pathtracer.cpp:17-28)
	if (rc != PT_OK) { fprintf(stderr, "%s: %s\n", what, pt_last_error()); exit(EXIT_FAILURE); }
}

Pathtracer::Pathtracer(uint32_t width, uint32_t height, unsigned int openglPixelBuffer)
	: m_width(width), m_height(height), m_hittableCount(0), m_nodeCount(0), m_timing(0.0f), m_accumulatedFrames(0)
{
	if (openglPixelBuffer != 0) { fprintf(stderr, "libpt_b200: OpenGL interop is not supported (headless)\n"); exit(EXIT_FAILURE); }
	pt_context *c = nullptr;
	check(pt_create(width, height, 0, &c), "pt_create");   // device 0, Pathtracer.cpp:40
	m_gpuAccumBuffer = reinterpret_cast<float4 *>(c);
}

Pathtracer::~Pathtracer() { pt_destroy(ctxOf(m_gpuAccumBuffer)); }

void Pathtracer::setScene(size_t count, const CpuHittable *hittables)   // Pathtracer.cpp:111-160
{
	std::vector<pt_object_xform_desc> d(count);
	for (size_t i = 0; i < count; ++i)
	{
		const Hittable h = hittables[i].getGpuHittable();
		unsigned char raw[sizeof(Hittable)];
		memcpy(raw, &h, sizeof raw);
		memcpy(d[i].world_to_local, raw, 48);
		memcpy(&d[i].type, raw + 88, 4);
		const unsigned char *m = raw + 48;
		memcpy(d[i].material.base_color, m, 12);
		memcpy(&d[i].material.roughness, m + 12, 4);
		memcpy(d[i].material.emissive, m + 16, 12);
		memcpy(&d[i].material.metalness, m + 28, 4);
		memcpy(&d[i].material.texture, m + 32, 4);
		memcpy(&d[i].material.type, m + 36, 4);
		const AABB &box = hittables[i].getAABB();
		d[i].aabb_min[0] = box.m_min.x; d[i].aabb_min[1] = box.m_min.y; d[i].aabb_min[2] = box.m_min.z;
		d[i].aabb_max[0] = box.m_max.x; d[i].aabb_max[1] = box.m_max.y; d[i].aabb_max[2] = box.m_max.z;
	}
	check(pt_set_scene_xform(ctxOf(m_gpuAccumBuffer), count, d.data()), "pt_set_scene_xform");
	m_hittableCount = uint32_t(count);
}

void Pathtracer::render(const Camera &camera, uint32_t spp, bool ignoreHistory)   // Pathtracer.cpp:162-227
{
	pt_camera_vectors v;
	const vec3 *src[4] = { &camera.m_origin, &camera.m_lowerLeftCorner, &camera.m_horizontal, &camera.m_vertical };
	float *dst[4] = { v.origin, v.lower_left, v.horizontal, v.vertical };
	for (int k = 0; k < 4; ++k) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
	check(pt_render_vectors(ctxOf(m_gpuAccumBuffer), &v, spp, ignoreHistory ? 1 : 0), "pt_render_vectors");
	m_timing = pt_get_timing_ms(ctxOf(m_gpuAccumBuffer));
}

float Pathtracer::getTiming() const { return m_timing; }                                                                   // :229-232
uint32_t Pathtracer::loadTexture(const char *path) { return pt_load_texture(ctxOf(m_gpuAccumBuffer), path); }              // :234-292
void Pathtracer::setSkyboxTextureHandle(uint32_t handle) { check(pt_set_skybox(ctxOf(m_gpuAccumBuffer), handle), "pt_set_skybox"); } // :294-297
float *Pathtracer::getHDRImageData() { return const_cast<float *>(pt_get_hdr(ctxOf(m_gpuAccumBuffer))); }                  // :299-315 (Q1 normalisation inside)
char *Pathtracer::getImageData() { return reinterpret_cast<char *>(const_cast<uint8_t *>(pt_get_ldr(ctxOf(m_gpuAccumBuffer)))); } // :317-339
