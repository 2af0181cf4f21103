// Microbenchmark: does sm_100's packed FFMA2 (fma.rn.f32x2) save ISSUE slots?  Four loops of independent chains, timed in SM
// clocks by one warp per scheduler (4 warps per SM, 1 CTA per SM) and by 8 warps per scheduler (the trace kernel's occupancy):
//   ffma        8 scalar FFMA chains
//   ffma2       8 packed FFMA2 chains (16 FMAs per 8 instructions)
//   ffma+lop    8 FFMA + 8 LOP3 (the ALU pipe) interleaved
//   ffma2+lop   8 FFMA2 + 8 LOP3 interleaved
// Output: warp-instructions per clock per SM sub-partition.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue ffma2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
constexpr int ITERS = 2048;
template <int MODE>
__global__ void k(float *out, long long *clocks, float a, float b)
{
	float f[8]; u64 p[8]; unsigned u[8];
	for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x + i; p[i] = (u64(__float_as_uint(f[i])) << 32) | __float_as_uint(f[i] + 1.0f); u[i] = threadIdx.x * 7 + i; }
	const u64 pa = (u64(__float_as_uint(a)) << 32) | __float_as_uint(a), pb = (u64(__float_as_uint(b)) << 32) | __float_as_uint(b);
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; ++it)
	{
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			if (MODE == 0 || MODE == 2) f[i] = fma1(f[i], a, b);
			if (MODE == 1 || MODE == 3) p[i] = fma2(p[i], pa, pb);
			if (MODE >= 2) u[i] = lop(u[i], u[(i + 1) & 7], 0x9e3779b9u);
		}
	}
	const long long t1 = clock64();
	float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(unsigned(p[i])) + __uint_as_float(unsigned(p[i] >> 32)) + float(u[i]);
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int threads, int instrPerIter)
{
	float *out; long long *clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
	k<MODE><<<148, threads>>>(out, clk, 1.0001f, 0.5f); cudaDeviceSynchronize();
	k<MODE><<<148, threads>>>(out, clk, 1.0001f, 0.5f); cudaDeviceSynchronize();
	long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < 148; ++i) avg += double(h[i]) / 148;
	const double warpsPerSmsp = threads / 32 / 4.0;
	printf("%-10s %4d threads/SM  %8.0f clocks  warp-instr/clk/SMSP %.3f\n", name, threads, avg, ITERS * double(instrPerIter) * warpsPerSmsp / avg);
	cudaFree(out); cudaFree(clk);
}
int main()
{
	for (int threads : { 128, 256, 1024 })
	{
		run<0>("ffma", threads, 8); run<1>("ffma2", threads, 8); run<2>("ffma+lop", threads, 16); run<3>("ffma2+lop", threads, 16);
	}
	return 0;
}
