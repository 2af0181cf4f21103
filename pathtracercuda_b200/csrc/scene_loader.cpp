// scene_loader.cpp — JSON scene -> object descriptions, field for field what the reference's loadScene does
// (SceneLoader.cpp:124-348), including its quirks:
//   * scalars (roughness, metalness, fovy) are read only from FLOAT literals (:163-171, Q5); vec3 elements accept ints
//   * unknown object/material type strings print a message and keep the default (:245-279, :290-308)
//   * rotation is degrees -> radians in float arithmetic: deg * (1.0f/180.0f) * pi (:193-196, :321)
//   * textures are de-duplicated by path string; "" means none (:127-149); the skybox shares the same table (:327-332)
#include "scene_loader.h"
#include <algorithm>
#include "json_min.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace ptb
{
namespace
{
float radiansf(float degree) { return degree * (1.0f / 180.0f) * 3.14159265358979323846f; }

bool getString(const JsonValue &o, const char *key, std::string &out)
{
	const JsonValue *v = o.find(key);
	if (v && v->kind == JsonValue::String) { out = v->str(); return true; }
	return false;
}
bool getFloat(const JsonValue &o, const char *key, float &out)
{
	const JsonValue *v = o.find(key);
	if (v && v->kind == JsonValue::Float) { out = float(v->num); return true; }
	return false;
}
// returns false on a type error inside the array (nlohmann would throw type_error.302 -> the reference terminates)
bool getVec3(const JsonValue &o, const char *key, float out[3], std::string &err)
{
	const JsonValue *v = o.find(key);
	if (v && v->kind == JsonValue::Array && v->arr().size() == 3)
	{
		for (int k = 0; k < 3; ++k)
		{
			const JsonValue &e = v->arr()[k];
			if (e.isNumber()) out[k] = float(e.num);
			else if (e.kind == JsonValue::Bool) out[k] = e.b ? 1.0f : 0.0f; // nlohmann converts booleans to numbers
			else { err = std::string("[json.exception.type_error.302] type must be number in \"") + key + "\""; return false; }
		}
	}
	return true;
}
} // namespace

bool parseSceneText(const std::string &text, float aspect, ParsedScene &out, std::string &err)
{
	out = ParsedScene();
	JsonValue root;
	if (!parseJson(text, root, err)) return false;

	auto textureHandle = [&](const std::string &p) -> uint32_t
	{
		if (p.empty()) return 0;
		for (size_t i = 0; i < out.texturePaths.size(); ++i)
			if (out.texturePaths[i] == p) return uint32_t(i + 1);
		out.texturePaths.push_back(p);
		return uint32_t(out.texturePaths.size());
	};

	const JsonValue *objs = root.find("objects");
	if (objs && objs->kind == JsonValue::Array)
	{
		out.hasObjectsArray = true;
		const std::vector<JsonValue> &list = objs->arr();
		const long count = long(list.size());
		const size_t ucount = list.size();
		out.objects.resize(ucount);
		// per object, by all cores: everything except the two things that depend on the order of the objects (texture handles
		// are given out in first-use order, the loader's messages are printed in object order) - those follow in one ordered pass
		std::vector<std::string> texOf(ucount);
		std::vector<std::vector<std::string>> msgOf(ucount);
		long firstBad = count;
		std::string firstErr;
#pragma omp parallel for schedule(static) if (count > 4096)
		for (long i = 0; i < count; ++i)
		{
			const JsonValue &o = list[size_t(i)];
			pt_object_desc d;
			memset(&d, 0, sizeof d);
			d.type = PT_SPHERE;
			d.scale[0] = d.scale[1] = d.scale[2] = 1.0f;
			d.material.type = PT_LAMBERT;
			d.material.base_color[0] = d.material.base_color[1] = d.material.base_color[2] = 1.0f;
			d.material.roughness = 0.5f;
			float rotationDeg[3] = { 0.0f, 0.0f, 0.0f };
			std::string e;

			std::string t;
			if (getString(o, "type", t))
			{
				static const char *names[] = { "SPHERE", "CYLINDER", "DISK", "CONE", "PARABOLOID", "QUAD", "CUBE" };
				bool found = false;
				for (uint32_t k = 0; k < 7; ++k) if (t == names[k]) { d.type = k; found = true; }
				if (!found) msgOf[size_t(i)].push_back("Failed to parse object type: " + t);
			}
			bool ok = getVec3(o, "position", d.position, e) && getVec3(o, "rotation", rotationDeg, e) && getVec3(o, "scale", d.scale, e);

			const JsonValue *m = ok ? o.find("material") : nullptr;
			if (m && m->kind == JsonValue::Object)
			{
				std::string mt;
				if (getString(*m, "type", mt))
				{
					if (mt == "LAMBERT") d.material.type = PT_LAMBERT;
					else if (mt == "GGX") d.material.type = PT_GGX;
					else if (mt == "LAMBERT_GGX") d.material.type = PT_LAMBERT_GGX;
					else msgOf[size_t(i)].push_back("Failed to parse material type: " + mt);
				}
				ok = getVec3(*m, "baseColor", d.material.base_color, e) && getVec3(*m, "emissive", d.material.emissive, e);
				getFloat(*m, "roughness", d.material.roughness);
				getFloat(*m, "metalness", d.material.metalness);
				getString(*m, "texture", texOf[size_t(i)]);
			}
			for (int k = 0; k < 3; ++k) d.rotation[k] = radiansf(rotationDeg[k]);
			out.objects[size_t(i)] = d;
			if (!ok)
			{
#pragma omp critical(ptb_scene_loader_error)
				if (i < firstBad) { firstBad = i; firstErr = e; }
			}
		}
		// the ordered pass; a type error ends the load at the first bad object, with the messages printed up to there
		for (long i = 0; i < std::min(count, firstBad + 1); ++i)
		{
			for (std::string &msg : msgOf[size_t(i)]) out.messages.push_back(std::move(msg));
			if (i == firstBad) { err = firstErr; return false; }
			if (!texOf[size_t(i)].empty()) out.objects[size_t(i)].material.texture = textureHandle(texOf[size_t(i)]);
		}
		// give the DOM of a large list back in parallel too (tens of millions of small nodes)
		if (count > 4096)
		{
			std::vector<JsonValue> &mut = const_cast<JsonValue *>(objs)->arr();
#pragma omp parallel for schedule(static)
			for (long i = 0; i < count; ++i) mut[size_t(i)] = JsonValue();
		}
	}

	std::string sky;
	if (getString(root, "skybox", sky)) { out.hasSkyboxString = true; out.skyboxTexture = textureHandle(sky); }

	float position[3] = { 0.0f, 0.0f, 0.0f }, lookAt[3] = { 0.0f, 0.0f, -1.0f }, fovy = 60.0f;
	const JsonValue *c = root.find("camera");
	if (c && c->kind == JsonValue::Object)
	{
		if (!getVec3(*c, "position", position, err) || !getVec3(*c, "look_at", lookAt, err)) return false;
		getFloat(*c, "fovy", fovy);
	}
	memcpy(out.camera.position, position, 12);
	memcpy(out.camera.look_at, lookAt, 12);
	out.camera.up[0] = 0.0f; out.camera.up[1] = 1.0f; out.camera.up[2] = 0.0f;
	out.camera.fovy = radiansf(fovy);
	out.camera.aspect = aspect;
	return true;
}

bool parseSceneFile(const char *path, float aspect, ParsedScene &out, std::string &err, int *errCode)
{
	std::ifstream f(path, std::ios::binary);
	if (!f.is_open())
	{
		err = std::string("Failed to open input file: ") + path; // SceneLoader.cpp:216
		if (errCode) *errCode = PT_E_IO;
		return false;
	}
	// one read into one buffer (a million-object scene file is hundreds of megabytes)
	std::string text;
	f.seekg(0, std::ios::end);
	const std::streamoff size = f.tellg();
	f.seekg(0, std::ios::beg);
	if (size > 0)
	{
		text.resize(size_t(size));
		f.read(&text[0], size);
		text.resize(size_t(f.gcount()));
	}
	if (!parseSceneText(text, aspect, out, err))
	{
		if (errCode) *errCode = PT_E_PARSE;
		return false;
	}
	return true;
}
} // namespace ptb
