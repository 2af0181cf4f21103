// Host-only stand-in for the reference's `class Pathtracer` member functions (declared in the reference's
// pathtracer/Pathtracer.h:12-69; the real definitions in pathtracer/Pathtracer.cpp need a GPU).  It lets the
// UNMODIFIED reference SceneLoader.cpp run on the CPU: loadScene() calls loadTexture / setScene /
// setSkyboxTextureHandle, and this mock simply records what it was given.  Harness file, ours; TEST
// INFRASTRUCTURE ONLY (built into oracle/_ref/libref_host.so).
#include "pathtracer/Pathtracer.h"
#include "ref_capture.h"
#include <cstdio>
#include <cstring>

RefCapture g_refCapture;

extern "C" int stbi_info(char const *filename, int *x, int *y, int *comp);

Pathtracer::Pathtracer(uint32_t width, uint32_t height, unsigned int)
	: m_width(width), m_height(height), m_hittableCount(), m_nodeCount(), m_timing(), m_accumulatedFrames()
{
	g_refCapture = RefCapture();
}

Pathtracer::~Pathtracer() {}

void Pathtracer::setScene(size_t count, const CpuHittable *hittables)
{
	if (count == 0)
	{
		printf("Setting an empty scene is not allowed!\n");
		return;
	}
	g_refCapture.objects.assign(hittables, hittables + count);
}

void Pathtracer::render(const Camera &, uint32_t, bool) {}
float Pathtracer::getTiming() const { return 0.0f; }

uint32_t Pathtracer::loadTexture(const char *path)
{
	// same success criterion as the reference (Pathtracer.cpp:245-257): the file must decode with stb_image
	if (m_textureCount >= 64)
	{
		return 0;
	}
	int w, h, c;
	if (!stbi_info(path, &w, &h, &c))
	{
		return 0;
	}
	g_refCapture.texturePaths.push_back(path);
	return ++m_textureCount;
}

void Pathtracer::setSkyboxTextureHandle(uint32_t handle) { g_refCapture.skybox = handle; }
float *Pathtracer::getHDRImageData() { return nullptr; }
char *Pathtracer::getImageData() { return nullptr; }
