// ref_gpu — GPU harness around the reference's OWN device code (kernels/trace.cu is #included where it lies
// under /root/reference at build time; nothing is copied).  TEST INFRASTRUCTURE ONLY.  Two jobs the
// unmodified reference binary (ref_pt) cannot do:
//   primary : deterministic primary-ray pass — pixel-centre rays through the reference's hitBVH
//             (kernels/trace.cu:28-98), dumping scene-order hit index and t per pixel (the parity gate).
//   tonemap : the reference's tonemap kernel (kernels/tonemap.cu, linked from its own object file) over an accumulation
//             buffer read from a file: the byte-exact gate for our tonemapKernel.
//   count   : the reference traceKernel's loop (kernels/trace.cu:158-199, getColor :101-156) with one counter
//             added, to learn how many rays (hitBVH calls) the reference traces for a given scene/size/spp
//             with its own XORWOW seeds (kernels/initRandState.cu:16) in its own 8-spp slices (main.cpp:271-278).
//             Untimed; the timing of the reference comes from the unmodified ref_pt binary.
// Input: a binary scene blob written by the tests/bench (tools: pathtracercuda_b200.refblob):
//   u32 magic 'PTB2', u32 width, u32 height, u32 objectCount, pt_camera_desc, pt_object_desc[objectCount]
// Output (primary): i32 index[W*H], f32 t[W*H] appended in one file.  Output (count): text line on stdout.
#include <cfloat>
#include <cstring>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <unordered_map>
#include "pathtracer/kernels/trace.cu"
#include "pathtracer/kernels/tonemap.h"
#include "../../include/pt_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_) { fprintf(stderr, "CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void primaryKernel(uint32_t width, uint32_t height, uint32_t hittableCount, Hittable *world, uint32_t nodeCount, BVHNode *nodes,
	Camera camera, const int32_t *bvhToScene, int32_t *hitIndex, float *hitT)
{
	uint32_t x = threadIdx.x + blockIdx.x * blockDim.x;
	uint32_t y = threadIdx.y + blockIdx.y * blockDim.y;
	if (x >= width || y >= height) return;
	float u = (x + 0.5f) / float(width);
	float v = (y + 0.5f) / float(height);
	Ray r = camera.getRay(u, v);
	HitRecord rec;
	rec.m_material = nullptr;
	bool hit = hitBVH(hittableCount, world, nodeCount, nodes, r, 0.001f, FLT_MAX, rec);
	// hitBVH does not return the element index: recover it from the material pointer (HitRecord.h:12 points into
	// the Hittable array; Material sits at byte 48 of the 96-byte Hittable, Hittable.h:22-27)
	int32_t idx = -1;
	if (hit)
	{
		size_t off = (const char *)rec.m_material - (const char *)world;
		idx = bvhToScene[off / sizeof(Hittable)];
	}
	hitIndex[x + y * width] = idx;
	hitT[x + y * width] = hit ? rec.m_t : 0.0f;
}

// kernels/trace.cu:101-156 with a ray counter; shading identical (skybox omitted: it cannot change path length)
__device__ uint32_t countRays(const Ray &r, uint32_t hittableCount, Hittable *world, uint32_t nodeCount, BVHNode *nodes, curandState &randState)
{
	vec3 throughput = vec3(1.0f);
	Ray ray = r;
	uint32_t rays = 0;
	for (int iteration = 0; iteration < 5; ++iteration)
	{
		HitRecord rec;
		++rays;
		bool found = hitBVH(hittableCount, world, nodeCount, nodes, ray, 0.001f, FLT_MAX, rec);
		if (!found) break;
		Ray scattered;
		float pdf = 0.0f;
		vec3 attenuation = rec.m_material->sample(ray, rec, randState, scattered, pdf, nullptr);
		if (attenuation == vec3(0.0f) || pdf == 0.0f) break;
		throughput *= attenuation * abs(dot(scattered.m_dir, rec.m_normal)) / pdf;
		ray = scattered;
	}
	return rays;
}

__global__ void initRand(int width, int height, uint32_t seedBase, curandState *st)
{
	int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
	if (x >= width || y >= height) return;
	uint32_t i = x + y * width;
	curand_init(seedBase + i, 0, 0, &st[i]);
}

__global__ void countKernel(uint32_t width, uint32_t height, uint32_t spp, uint32_t hittableCount, Hittable *world, uint32_t nodeCount, BVHNode *nodes,
	curandState *randState, Camera camera, unsigned long long *total)
{
	int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + blockIdx.y * blockDim.y;
	if (x >= width || y >= height) return;
	uint32_t dstIdx = x + y * width;
	curandState &st = randState[dstIdx];
	unsigned long long rays = 0;
	for (uint32_t i = 0; i < spp; ++i)
	{
		float u = (x + curand_uniform(&st)) / float(width);
		float v = (y + curand_uniform(&st)) / float(height);
		Ray r = camera.getRay(u, v);
		rays += countRays(r, hittableCount, world, nodeCount, nodes, st);
	}
	atomicAdd(total, rays);
}

int main(int argc, char **argv)
{
	if (argc < 4)
	{
		fprintf(stderr, "usage: ref_gpu primary <scene.blob> <out.bin> | ref_gpu count <scene.blob> <spp>\n");
		return 2;
	}
	if (strcmp(argv[1], "tonemap") == 0)
	{
		// ref_gpu tonemap <accum.bin> <out.bin> <width> <height> <accumulatedSampleCount>: the reference's own tonemap kernel
		// (kernels/tonemap.cu:4-27, launched as in Pathtracer.cpp:322-328) over a float4 accumulation buffer from a file
		if (argc < 7) { fprintf(stderr, "usage: ref_gpu tonemap <accum.bin> <out.bin> <width> <height> <count>\n"); return 2; }
		const uint32_t W = (uint32_t)atoi(argv[4]), H = (uint32_t)atoi(argv[5]), count = (uint32_t)atoi(argv[6]);
		std::vector<float> acc(size_t(W) * H * 4);
		FILE *fi = fopen(argv[2], "rb");
		if (!fi || fread(acc.data(), 16, size_t(W) * H, fi) != size_t(W) * H) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
		fclose(fi);
		float4 *dAcc; uchar4 *dOut;
		CK(cudaMalloc(&dAcc, acc.size() * 4)); CK(cudaMalloc(&dOut, size_t(W) * H * 4));
		CK(cudaMemcpy(dAcc, acc.data(), acc.size() * 4, cudaMemcpyHostToDevice));
		dim3 threads(8, 8, 1), blocks((W + 7) / 8, (H + 7) / 8, 1);
		tonemap<<<blocks, threads>>>(dOut, dAcc, W, H, count);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		std::vector<unsigned char> out(size_t(W) * H * 4);
		CK(cudaMemcpy(out.data(), dOut, out.size(), cudaMemcpyDeviceToHost));
		FILE *fo = fopen(argv[3], "wb");
		if (!fo) { fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
		fwrite(out.data(), 1, out.size(), fo); fclose(fo);
		printf("{\"mode\":\"tonemap\",\"width\":%u,\"height\":%u,\"count\":%u}\n", W, H, count);
		return 0;
	}
	FILE *f = fopen(argv[2], "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
	uint32_t hdr[4];
	pt_camera_desc cd;
	if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 0x32425450u || fread(&cd, sizeof(cd), 1, f) != 1) { fprintf(stderr, "bad blob\n"); return 2; }
	uint32_t W = hdr[1], H = hdr[2], N = hdr[3];
	std::vector<pt_object_desc> descs(N);
	if (fread(descs.data(), sizeof(pt_object_desc), N, f) != N) { fprintf(stderr, "short blob\n"); return 2; }
	fclose(f);

	std::vector<CpuHittable> objects;
	for (auto &d : descs)
	{
		const pt_material_desc &m = d.material;
		Material mat((MaterialType)m.type, vec3(m.base_color[0], m.base_color[1], m.base_color[2]), vec3(m.emissive[0], m.emissive[1], m.emissive[2]), m.roughness, m.metalness, 0);
		objects.push_back(CpuHittable((HittableType)d.type, vec3(d.position[0], d.position[1], d.position[2]), vec3(d.rotation[0], d.rotation[1], d.rotation[2]), vec3(d.scale[0], d.scale[1], d.scale[2]), mat));
	}
	BVH bvh;
	bvh.build(objects.size(), objects.data(), 4);
	std::vector<Hittable> gpuH;
	std::vector<int32_t> bvhToScene(N, -1);
	std::vector<char> used(N, 0);
	const auto &elems = bvh.getElements();
	// BVH order -> scene order: BVH::build reorders copies of the CpuHittables; match them back by their bytes (hashed: the
	// synthetic scenes of BASELINE config 4 have up to a million objects), duplicates in scene order
	std::unordered_multimap<uint64_t, uint32_t> byHash;
	byHash.reserve(N * 2);
	auto hashOf = [](const CpuHittable &h) { uint64_t x = 1469598103934665603ull; const unsigned char *p = (const unsigned char *)&h; for (size_t k = 0; k < sizeof(CpuHittable); ++k) x = (x ^ p[k]) * 1099511628211ull; return x; };
	for (size_t j = N; j-- > 0;) byHash.emplace(hashOf(objects[j]), (uint32_t)j);
	for (size_t i = 0; i < elems.size(); ++i)
	{
		gpuH.push_back(elems[i].getGpuHittable());
		auto range = byHash.equal_range(hashOf(elems[i]));
		uint32_t bestJ = 0xffffffffu;
		for (auto it = range.first; it != range.second; ++it)
			if (!used[it->second] && it->second < bestJ && memcmp(&elems[i], &objects[it->second], sizeof(CpuHittable)) == 0) bestJ = it->second;
		if (bestJ != 0xffffffffu) { used[bestJ] = 1; bvhToScene[i] = (int32_t)bestJ; }
	}
	Camera camera(vec3(cd.position[0], cd.position[1], cd.position[2]), vec3(cd.look_at[0], cd.look_at[1], cd.look_at[2]), vec3(cd.up[0], cd.up[1], cd.up[2]), cd.fovy, cd.aspect);

	BVHNode *dNodes; Hittable *dH; int32_t *dMap;
	CK(cudaMalloc(&dNodes, bvh.getNodes().size() * sizeof(BVHNode)));
	CK(cudaMalloc(&dH, gpuH.size() * sizeof(Hittable)));
	CK(cudaMalloc(&dMap, N * sizeof(int32_t)));
	CK(cudaMemcpy(dNodes, bvh.getNodes().data(), bvh.getNodes().size() * sizeof(BVHNode), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dH, gpuH.data(), gpuH.size() * sizeof(Hittable), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dMap, bvhToScene.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
	dim3 threads(8, 8, 1), blocks((W + 7) / 8, (H + 7) / 8, 1);

	if (strcmp(argv[1], "primary") == 0)
	{
		int32_t *dIdx; float *dT;
		CK(cudaMalloc(&dIdx, size_t(W) * H * 4)); CK(cudaMalloc(&dT, size_t(W) * H * 4));
		primaryKernel<<<blocks, threads>>>(W, H, (uint32_t)gpuH.size(), dH, (uint32_t)bvh.getNodes().size(), dNodes, camera, dMap, dIdx, dT);
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		std::vector<int32_t> idx(size_t(W) * H); std::vector<float> t(size_t(W) * H);
		CK(cudaMemcpy(idx.data(), dIdx, idx.size() * 4, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(t.data(), dT, t.size() * 4, cudaMemcpyDeviceToHost));
		FILE *o = fopen(argv[3], "wb");
		if (!o) { fprintf(stderr, "cannot write %s\n", argv[3]); return 2; }
		fwrite(idx.data(), 4, idx.size(), o); fwrite(t.data(), 4, t.size(), o); fclose(o);
		printf("{\"mode\":\"primary\",\"width\":%u,\"height\":%u,\"objects\":%u,\"nodes\":%zu,\"depth\":%u}\n", W, H, N, bvh.getNodes().size(), bvh.getDepth());
	}
	else
	{
		uint32_t spp = (uint32_t)atoi(argv[3]);
		curandState *dSt; unsigned long long *dTotal;
		CK(cudaMalloc(&dSt, size_t(W) * H * sizeof(curandState))); CK(cudaMalloc(&dTotal, 8)); CK(cudaMemset(dTotal, 0, 8));
		initRand<<<blocks, threads>>>(W, H, 1984u, dSt);
		for (uint32_t i = 0; i < spp; i += 8)
		{
			uint32_t s = (i + 8 < spp ? i + 8 : spp) - i;
			countKernel<<<blocks, threads>>>(W, H, s, (uint32_t)gpuH.size(), dH, (uint32_t)bvh.getNodes().size(), dNodes, dSt, camera, dTotal);
		}
		CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
		unsigned long long total = 0;
		CK(cudaMemcpy(&total, dTotal, 8, cudaMemcpyDeviceToHost));
		printf("{\"mode\":\"count\",\"width\":%u,\"height\":%u,\"spp\":%u,\"objects\":%u,\"rays\":%llu,\"samples\":%llu,\"rays_per_sample\":%.6f}\n",
			W, H, spp, N, total, (unsigned long long)W * H * spp, double(total) / (double(W) * H * spp));
	}
	return 0;
}
