// libref_host.so — the reference's OWN __host__ __device__ math (vec3, AABB, Hittable::hit, Camera, BVH::build,
// Material::sample with cuRAND XORWOW, brdf.h, MonteCarlo.h) compiled for the CPU with g++, behind a small C
// interface.  TEST INFRASTRUCTURE ONLY: it pins oracle/pt_oracle.c (the CPU restatement) and is the
// "reference" CPU baseline of bench.py.  Nothing here is product code and nothing is copied from the
// reference: the reference headers are #included where they lie under /root/reference at build time
// (oracle/Makefile), only the two device-only drivers are restated, each citing the lines it follows:
//   hitBVH    <- kernels/trace.cu:28-98   (calls the reference AABB::hit / Hittable::hit)
//   getColor  <- kernels/trace.cu:101-156 (calls the reference Material::sample / getEmitted)
// The skybox tex2D (trace.cu:128) and the base-colour tex2D (Material.inl:26-35, compiled out on the host by
// the reference's own `#if __CUDA_ARCH__`) need the texture unit; the host build uses a software bilinear
// lookup for the sky (normalized coords, wrap U / clamp V, texel centres at +0.5 — the CUDA rule,
// Pathtracer.cpp:276-281) and ignores base-colour textures.
#include <cfloat>
#include <cstring>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include "pathtracer/BVH.h"
#include "pathtracer/Camera.h"
#include "pathtracer/Hittable.h"
#include "pathtracer/Material.h"
#include "pathtracer/Pathtracer.h"
#include "Params.h"
#include "SceneLoader.h"
#include "ref_capture.h"
#include "../../include/pt_b200.h"

namespace
{
	std::vector<CpuHittable> g_objects;   // scene order
	std::vector<Hittable> g_bvhHittables; // BVH order
	std::vector<BVHNode> g_nodes;
	std::vector<int32_t> g_bvhToScene;
	Camera g_camera(vec3(0.0f), vec3(0.0f, 0.0f, -1.0f), vec3(0.0f, 1.0f, 0.0f), 1.0f, 1.0f);
	std::vector<float> g_sky; // RGBA float
	int g_skyW = 0, g_skyH = 0;
	uint64_t g_nodeVisits = 0, g_primTests = 0;

	void buildBVH()
	{
		BVH bvh;
		bvh.build(g_objects.size(), g_objects.data(), 4); // Pathtracer.cpp:121
		g_nodes = bvh.getNodes();
		const auto &elems = bvh.getElements();
		g_bvhHittables.clear();
		g_bvhToScene.assign(elems.size(), -1);
		std::vector<char> used(g_objects.size(), 0);
		for (size_t i = 0; i < elems.size(); ++i)
		{
			g_bvhHittables.push_back(elems[i].getGpuHittable()); // Pathtracer.cpp:125-133
			for (size_t j = 0; j < g_objects.size(); ++j)
			{
				if (!used[j] && memcmp(&elems[i], &g_objects[j], sizeof(CpuHittable)) == 0)
				{
					used[j] = 1;
					g_bvhToScene[i] = (int32_t)j;
					break;
				}
			}
		}
	}

	// restatement of kernels/trace.cu:28-98 (device-only in the reference); returns BVH-order element index
	bool hitBVH(const Ray &r, float t_min, float t_max, HitRecord &rec, uint32_t &elemIdxOut, uint64_t *nodeVisits, uint64_t *primTests)
	{
		vec3 invRayDir;
		invRayDir[0] = 1.0f / (r.m_dir[0] != 0.0f ? r.m_dir[0] : 1e-7f);
		invRayDir[1] = 1.0f / (r.m_dir[1] != 0.0f ? r.m_dir[1] : 1e-7f);
		invRayDir[2] = 1.0f / (r.m_dir[2] != 0.0f ? r.m_dir[2] : 1e-7f);
		bool dirIsNeg[3] = { invRayDir.x < 0.0f, invRayDir.y < 0.0f, invRayDir.z < 0.0f };
		uint32_t nodesToVisit[64];
		uint32_t toVisitOffset = 0;
		uint32_t currentNodeIndex = 0;
		uint32_t elemIdx = UINT32_MAX;
		const BVHNode *bvhNodes = g_nodes.data();
		const Hittable *world = g_bvhHittables.data();
		while (true)
		{
			const BVHNode &node = bvhNodes[currentNodeIndex];
			if (nodeVisits) ++*nodeVisits;
			if (node.m_aabb.hit(r, t_min, t_max))
			{
				const uint32_t primitiveCount = (node.m_primitiveCountAxis >> 16);
				if (primitiveCount > 0)
				{
					for (uint32_t i = 0; i < primitiveCount; ++i)
					{
						if (primTests) ++*primTests;
						if (world[node.m_offset + i].hit(r, t_min, t_max, rec))
						{
							t_max = rec.m_t;
							elemIdx = node.m_offset + i;
						}
					}
					if (toVisitOffset == 0) break;
					currentNodeIndex = nodesToVisit[--toVisitOffset];
				}
				else
				{
					bool isNeg = dirIsNeg[(node.m_primitiveCountAxis >> 8) & 0xFF];
					nodesToVisit[toVisitOffset++] = isNeg ? (currentNodeIndex + 1) : node.m_offset;
					currentNodeIndex = isNeg ? node.m_offset : (currentNodeIndex + 1);
				}
			}
			else
			{
				if (toVisitOffset == 0) break;
				currentNodeIndex = nodesToVisit[--toVisitOffset];
			}
		}
		elemIdxOut = elemIdx;
		return elemIdx != UINT32_MAX;
	}

	vec3 skyLookup(float u, float v)
	{
		// CUDA linear filtering, normalized coordinates: x = u*W - 0.5; wrap in U, clamp in V
		float x = u * g_skyW - 0.5f, y = v * g_skyH - 0.5f;
		float fx = floorf(x), fy = floorf(y);
		float ax = x - fx, ay = y - fy;
		int x0 = (int)fx, y0 = (int)fy;
		auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
		auto clampi = [](int i, int n) { return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); };
		int xa = wrap(x0, g_skyW), xb = wrap(x0 + 1, g_skyW), ya = clampi(y0, g_skyH), yb = clampi(y0 + 1, g_skyH);
		vec3 r(0.0f);
		for (int c = 0; c < 3; ++c)
		{
			float t00 = g_sky[(size_t(ya) * g_skyW + xa) * 4 + c], t10 = g_sky[(size_t(ya) * g_skyW + xb) * 4 + c];
			float t01 = g_sky[(size_t(yb) * g_skyW + xa) * 4 + c], t11 = g_sky[(size_t(yb) * g_skyW + xb) * 4 + c];
			r[c] = (1.0f - ay) * ((1.0f - ax) * t00 + ax * t10) + ay * ((1.0f - ax) * t01 + ax * t11);
		}
		return r;
	}

	// restatement of kernels/trace.cu:101-156
	vec3 getColor(const Ray &r, curandState &randState, uint64_t *rays)
	{
		vec3 throughput = vec3(1.0f);
		vec3 L = vec3(0.0f);
		Ray ray = r;
		for (int iteration = 0; iteration < 5; ++iteration)
		{
			HitRecord rec;
			uint32_t idx;
			++*rays;
			bool foundIntersection = hitBVH(ray, 0.001f, FLT_MAX, rec, idx, nullptr, nullptr);
			if (!foundIntersection)
			{
				vec3 c = 0.0f;
				if (g_skyW > 0)
				{
					float theta = acosf(ray.m_dir.y);
					float phi = atan2f(ray.m_dir.z, ray.m_dir.x);
					float v = theta / PI;
					float u = phi / (2.0f * PI);
					c = skyLookup(u, v);
				}
				L += throughput * c;
				break;
			}
			else
			{
				L += throughput * rec.m_material->getEmitted(ray, rec);
				Ray scattered;
				float pdf = 0.0f;
				vec3 attenuation = rec.m_material->sample(ray, rec, randState, scattered, pdf, nullptr);
				if (attenuation == vec3(0.0f) || pdf == 0.0f)
				{
					break;
				}
				throughput *= attenuation * fabsf(dot(scattered.m_dir, rec.m_normal)) / pdf;
				ray = scattered;
			}
		}
		return L;
	}

	Camera makeCamera(const pt_camera_desc *c)
	{
		return Camera(vec3(c->position[0], c->position[1], c->position[2]), vec3(c->look_at[0], c->look_at[1], c->look_at[2]),
			vec3(c->up[0], c->up[1], c->up[2]), c->fovy, c->aspect);
	}
}

#include <omp.h>
extern "C"
{
	// --- scene input -------------------------------------------------------------------------------------
	// through the reference's own loadScene (SceneLoader.cpp:124-348); texture paths resolve against the CWD
	int refh_load_scene_file(const char *path, uint32_t width, uint32_t height)
	{
		Params params;
		params.m_width = width;
		params.m_height = height;
		params.m_inputFilepath = path;
		Pathtracer pt(width, height, 0);
		g_camera = loadScene(pt, params);
		g_objects = g_refCapture.objects;
		if (!g_objects.empty()) buildBVH();
		return (int)g_objects.size();
	}

	// through the reference's CpuHittable / Material ctors (Hittable.cpp:115-179, Material.inl:8-18)
	int refh_set_scene(size_t n, const pt_object_desc *d)
	{
		g_objects.clear();
		for (size_t i = 0; i < n; ++i)
		{
			const pt_material_desc &m = d[i].material;
			Material mat((MaterialType)m.type, vec3(m.base_color[0], m.base_color[1], m.base_color[2]),
				vec3(m.emissive[0], m.emissive[1], m.emissive[2]), m.roughness, m.metalness, m.texture);
			g_objects.push_back(CpuHittable((HittableType)d[i].type, vec3(d[i].position[0], d[i].position[1], d[i].position[2]),
				vec3(d[i].rotation[0], d[i].rotation[1], d[i].rotation[2]), vec3(d[i].scale[0], d[i].scale[1], d[i].scale[2]), mat));
		}
		if (!g_objects.empty()) buildBVH();
		return (int)g_objects.size();
	}

	void refh_set_camera(const pt_camera_desc *c) { g_camera = makeCamera(c); }
	void refh_get_camera(float *out23)
	{
		// m_tanHalfFovy, m_aspectRatio, origin, lowerLeft, horizontal, vertical, right, up, backward
		memcpy(out23, &g_camera, sizeof(float) * 23);
	}
	// the reference's own interactive camera controls (Camera.inl:30-52) on the current camera
	void refh_camera_rotate(float pitch, float yaw, float roll) { g_camera.rotate(pitch, yaw, roll); }
	void refh_camera_translate(float x, float y, float z) { g_camera.translate(x, y, z); }
	int refh_texture_count() { return (int)g_refCapture.texturePaths.size(); }
	const char *refh_texture_path(int i) { return g_refCapture.texturePaths[i].c_str(); }
	uint32_t refh_skybox_handle() { return g_refCapture.skybox; }
	void refh_set_sky(int w, int h, const float *rgba)
	{
		g_skyW = w;
		g_skyH = h;
		g_sky.assign(rgba, rgba + size_t(w) * h * 4);
	}

	// --- inspection --------------------------------------------------------------------------------------
	// raw bytes of object i in scene order: 3 x float4 world->local rows, Material (40 B), AABB (24 B), type (4 B)
	int refh_object_bytes(int i, void *out128)
	{
		if (i < 0 || (size_t)i >= g_objects.size()) return -1;
		memcpy(out128, &g_objects[i], sizeof(CpuHittable));
		return (int)sizeof(CpuHittable);
	}
	int refh_bvh_info(uint32_t *nodeCount, uint32_t *depth, int32_t *valid)
	{
		BVH bvh;
		bvh.build(g_objects.size(), g_objects.data(), 4);
		*nodeCount = (uint32_t)bvh.getNodes().size();
		*depth = bvh.getDepth();
		*valid = bvh.validate() ? 1 : 0;
		return 0;
	}
	int refh_bvh_nodes(void *out, size_t capacityNodes)
	{
		size_t n = std::min(capacityNodes, g_nodes.size());
		memcpy(out, g_nodes.data(), n * sizeof(BVHNode));
		return (int)g_nodes.size();
	}
	void refh_bvh_order(int32_t *bvhToScene) { memcpy(bvhToScene, g_bvhToScene.data(), g_bvhToScene.size() * sizeof(int32_t)); }

	// --- unit probes (reference functions, host build) -----------------------------------------------------
	// Hittable::hit of scene object i (Hittable.inl:88-145); out = t, p[3], n[3], u, v, frontFace
	int refh_hit_object(int i, const float *o, const float *d, float tMin, float tMax, float *out10)
	{
		Hittable h = g_objects[i].getGpuHittable();
		HitRecord rec;
		rec.m_texCoordU = 0.0f;
		rec.m_texCoordV = 0.0f;
		Ray r(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]));
		if (!h.hit(r, tMin, tMax, rec)) return 0;
		out10[0] = rec.m_t;
		out10[1] = rec.m_p.x; out10[2] = rec.m_p.y; out10[3] = rec.m_p.z;
		out10[4] = rec.m_normal.x; out10[5] = rec.m_normal.y; out10[6] = rec.m_normal.z;
		out10[7] = rec.m_texCoordU; out10[8] = rec.m_texCoordV; out10[9] = rec.m_frontFace ? 1.0f : 0.0f;
		return 1;
	}
	int refh_aabb_hit(const float *mn, const float *mx, const float *o, const float *d, float tMin, float tMax)
	{
		AABB b(vec3(mn[0], mn[1], mn[2]), vec3(mx[0], mx[1], mx[2]));
		return b.hit(Ray(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2])), tMin, tMax) ? 1 : 0;
	}
	void refh_camera_ray(float s, float t, float *out6)
	{
		Ray r = g_camera.getRay(s, t);
		out6[0] = r.m_origin.x; out6[1] = r.m_origin.y; out6[2] = r.m_origin.z;
		out6[3] = r.m_dir.x; out6[4] = r.m_dir.y; out6[5] = r.m_dir.z;
	}
	// Material::sample (Material.inl:20-60) with the two uniforms it will draw reported back.
	// in: material desc, world normal N (unit, already face-forwarded), incoming ray direction; xorwow seed.
	// out13 = rnd0, rnd1, attenuation[3], pdf, scattered dir[3], emitted[3], spare
	void refh_material_sample(const pt_material_desc *m, const float *N, const float *inDir, uint32_t seed, float *out13)
	{
		Material mat((MaterialType)m->type, vec3(m->base_color[0], m->base_color[1], m->base_color[2]),
			vec3(m->emissive[0], m->emissive[1], m->emissive[2]), m->roughness, m->metalness, 0);
		curandState st;
		curand_init(seed, 0, 0, &st);
		curandState peek = st;
		out13[0] = curand_uniform(&peek);
		out13[1] = curand_uniform(&peek);
		HitRecord rec;
		rec.m_p = vec3(0.0f);
		rec.m_normal = vec3(N[0], N[1], N[2]);
		rec.m_t = 1.0f;
		rec.m_material = &mat;
		rec.m_texCoordU = rec.m_texCoordV = 0.0f;
		rec.m_frontFace = true;
		Ray rin(vec3(0.0f), vec3(inDir[0], inDir[1], inDir[2]));
		Ray sc;
		float pdf = 0.0f;
		vec3 att = mat.sample(rin, rec, st, sc, pdf, nullptr);
		vec3 em = mat.getEmitted(rin, rec);
		out13[2] = att.x; out13[3] = att.y; out13[4] = att.z; out13[5] = pdf;
		out13[6] = sc.m_dir.x; out13[7] = sc.m_dir.y; out13[8] = sc.m_dir.z;
		out13[9] = em.x; out13[10] = em.y; out13[11] = em.z; out13[12] = 0.0f;
	}
	// kernels/tonemap.cu:4-27 arithmetic on the host (powf)
	void refh_tonemap(const float *accumRGBA, size_t pixels, uint32_t sampleCount, uint8_t *outRGBA)
	{
		for (size_t i = 0; i < pixels; ++i)
		{
			vec3 c = vec3(accumRGBA[i * 4 + 0], accumRGBA[i * 4 + 1], accumRGBA[i * 4 + 2]) / float(sampleCount);
			c = c / (c + 1.0f);
			c.r = powf(c.r, 1.0f / 2.2f); c.g = powf(c.g, 1.0f / 2.2f); c.b = powf(c.b, 1.0f / 2.2f);
			outRGBA[i * 4 + 0] = (unsigned char)(c.x * 255.0f);
			outRGBA[i * 4 + 1] = (unsigned char)(c.y * 255.0f);
			outRGBA[i * 4 + 2] = (unsigned char)(c.z * 255.0f);
			outRGBA[i * 4 + 3] = 255;
		}
	}

	// --- passes ------------------------------------------------------------------------------------------
	// deterministic primary pass (SURVEY.md §8d parity gate): pixel-centre rays, t_min = 0.001, scene-order index
	void refh_primary_pass(uint32_t width, uint32_t height, int32_t *hitIndex, float *hitT, uint64_t *stats2)
	{
		uint64_t nv = 0, np = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : nv, np)
		for (int y = 0; y < (int)height; ++y)
		{
			Camera cam = g_camera;
			for (uint32_t x = 0; x < width; ++x)
			{
				float u = (x + 0.5f) / float(width);
				float v = (y + 0.5f) / float(height);
				Ray r = cam.getRay(u, v);
				HitRecord rec;
				uint32_t idx;
				bool hit = !g_nodes.empty() && hitBVH(r, 0.001f, FLT_MAX, rec, idx, &nv, &np);
				hitIndex[size_t(y) * width + x] = hit ? g_bvhToScene[idx] : -1;
				hitT[size_t(y) * width + x] = hit ? rec.m_t : 0.0f;
			}
		}
		if (stats2) { stats2[0] = nv; stats2[1] = np; }
	}

	// closest hit for caller rays; normal out optional (3 floats per ray)
	void refh_trace_rays(size_t n, const float *o, const float *d, float tMin, int32_t *hitIndex, float *hitT, float *hitN)
	{
#pragma omp parallel for schedule(static, 256)
		for (long i = 0; i < (long)n; ++i)
		{
			Ray r(vec3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), vec3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
			HitRecord rec;
			uint32_t idx;
			bool hit = hitBVH(r, tMin, FLT_MAX, rec, idx, nullptr, nullptr);
			hitIndex[i] = hit ? g_bvhToScene[idx] : -1;
			hitT[i] = hit ? rec.m_t : 0.0f;
			if (hitN) { hitN[3 * i] = hit ? rec.m_normal.x : 0.0f; hitN[3 * i + 1] = hit ? rec.m_normal.y : 0.0f; hitN[3 * i + 2] = hit ? rec.m_normal.z : 0.0f; }
		}
	}

	// full path tracing on the host following kernels/trace.cu:158-199 (XORWOW seed = seedBase + pixel, as
	// kernels/initRandState.cu:16).  accum = SUM over spp samples (RGBA, A = 1).  Returns rays traced.
	// threads the next refh_render will use (bench.py reports them as cpu_baseline.cores); n > 0 sets the count first -
	// torchrun exports OMP_NUM_THREADS=1 to its workers, which would time one core and call it the box
	int refh_threads(int n)
	{
		if (n > 0) omp_set_num_threads(n);
		return omp_get_max_threads();
	}
	uint64_t refh_render(uint32_t width, uint32_t height, uint32_t spp, uint32_t seedBase, float *accumRGBA)
	{
		uint64_t rays = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays)
		for (int y = 0; y < (int)height; ++y)
		{
			Camera cam = g_camera;
			for (uint32_t x = 0; x < width; ++x)
			{
				uint32_t dstIdx = x + y * width;
				curandState st;
				curand_init(seedBase + dstIdx, 0, 0, &st);
				vec3 color = 0.0f;
				for (uint32_t i = 0; i < spp; ++i)
				{
					float u = (x + curand_uniform(&st)) / float(width);
					float v = (y + curand_uniform(&st)) / float(height);
					Ray r = cam.getRay(u, v);
					color += getColor(r, st, &rays);
				}
				accumRGBA[dstIdx * 4 + 0] = color.r; accumRGBA[dstIdx * 4 + 1] = color.g;
				accumRGBA[dstIdx * 4 + 2] = color.b; accumRGBA[dstIdx * 4 + 3] = 1.0f;
			}
		}
		return rays;
	}
}
