// What the mock Pathtracer records from the reference's loadScene().  Harness file, ours.
#pragma once
#include "pathtracer/Hittable.h"
#include <string>
#include <vector>
struct RefCapture
{
	std::vector<CpuHittable> objects;
	std::vector<std::string> texturePaths;
	uint32_t skybox = 0;
};
extern RefCapture g_refCapture;
