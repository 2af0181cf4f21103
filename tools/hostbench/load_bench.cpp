// Host-only timing of the scene load path (no GPU): file -> objects (scene_loader) -> BVH + records (scene_compile).
// build: g++ -O3 -fopenmp -ffp-contract=off -std=c++17 -I../../pathtracercuda_b200/csrc load_bench.cpp ../../pathtracercuda_b200/csrc/{scene_loader,scene_compile,json_min}.cpp -o /tmp/load_bench
#include "scene_compile.h"
#include "scene_loader.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <omp.h>
using namespace ptb;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
	if (argc < 2) return 1;
	for (int rep = 0; rep < 3; ++rep)
	{
		ParsedScene ps;
		std::string err;
		int code = 0;
		const double t0 = now();
		if (!parseSceneFile(argv[1], 16.0f / 9.0f, ps, err, &code)) { printf("parse failed: %s\n", err.c_str()); return 1; }
		const double t1 = now();
		CompiledScene cs;
		if (!compileScene(ps.objects.size(), ps.objects.data(), 4, cs, err, kMaxGlobalPrims, nullptr, argc > 2 ? atoi(argv[2]) : 0)) { printf("compile failed: %s\n", err.c_str()); return 1; }
		const double t2 = now();
		printf("threads %d objects %zu: parse %.3f s, compile %.3f s (nodes %zu, depth %u)\n", omp_get_max_threads(), ps.objects.size(), t1 - t0, t2 - t1, cs.nodes.size(), cs.depth);
	}
	return 0;
}
