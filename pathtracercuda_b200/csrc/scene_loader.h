// scene_loader.h — the scene JSON schema of the reference (loadScene, SceneLoader.cpp:124-348) as plain data.
#pragma once
#include "../../include/pt_b200.h"
#include <string>
#include <vector>

namespace ptb
{
struct ParsedScene
{
	std::vector<pt_object_desc> objects;     // material.texture = 1-based index into texturePaths (0 = none)
	std::vector<std::string> texturePaths;   // unique, in first-use order (objects first, then the skybox)
	uint32_t skyboxTexture = 0;              // 1-based index into texturePaths, 0 = none
	bool hasObjectsArray = false;
	bool hasSkyboxString = false;            // "skybox" is a string: the loader calls setSkyboxTextureHandle (with 0 for "" or a failed load)
	pt_camera_desc camera;
	std::vector<std::string> messages;       // the loader's printf lines ("Failed to parse object type: ...")
};

bool parseSceneFile(const char *path, float aspect, ParsedScene &out, std::string &err, int *errCode);
bool parseSceneText(const std::string &text, float aspect, ParsedScene &out, std::string &err);
bool parseSceneText(const char *text, size_t n, float aspect, ParsedScene &out, std::string &err); // (need not be null-terminated)
} // namespace ptb
