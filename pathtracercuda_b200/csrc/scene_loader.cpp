// scene_loader.cpp — JSON scene -> object descriptions, field for field what the reference's loadScene does
// (SceneLoader.cpp:124-348), including its quirks:
//   * scalars (roughness, metalness, fovy) are read only from FLOAT literals (:163-171, Q5); vec3 elements accept ints
//   * unknown object/material type strings print a message and keep the default (:245-279, :290-308)
//   * rotation is degrees -> radians in float arithmetic: deg * (1.0f/180.0f) * pi (:193-196, :321)
//   * textures are de-duplicated by path string; "" means none (:127-149); the skybox shares the same table (:327-332)
#include "scene_loader.h"
#include <algorithm>
#include <climits>
#include "json_min.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include <chrono>
#include <cstdlib>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
namespace ptb
{
namespace
{
// PTB_TIMING=1: phase timings of the loader on stderr
struct LoaderLap
{
	bool on = getenv("PTB_TIMING") != nullptr;
	double last = now();
	static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	void operator()(const char *what) { if (on) { const double t = now(); fprintf(stderr, "  loadScene    %-14s %7.1f ms\n", what, (t - last) * 1e3); last = t; } }
};
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

// ---- one element of "objects": DOM value -> description, field for field what loadScene does (SceneLoader.cpp:228-322) ----
void objectDefaults(pt_object_desc &d)
{
	memset(&d, 0, sizeof d);
	d.type = PT_SPHERE;
	d.scale[0] = d.scale[1] = d.scale[2] = 1.0f;
	d.material.type = PT_LAMBERT;
	d.material.base_color[0] = d.material.base_color[1] = d.material.base_color[2] = 1.0f;
	d.material.roughness = 0.5f;
}
const char *const kShapeNames[7] = { "SPHERE", "CYLINDER", "DISK", "CONE", "PARABOLOID", "QUAD", "CUBE" };

bool convertObject(const JsonValue &o, pt_object_desc &d, std::string &tex, std::vector<std::string> &msgs, std::string &e)
{
	objectDefaults(d);
	float rotationDeg[3] = { 0.0f, 0.0f, 0.0f };
	std::string t;
	if (getString(o, "type", t))
	{
		bool found = false;
		for (uint32_t k = 0; k < 7; ++k) if (t == kShapeNames[k]) { d.type = k; found = true; }
		if (!found) msgs.push_back("Failed to parse object type: " + t);
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
			else msgs.push_back("Failed to parse material type: " + mt);
		}
		ok = getVec3(*m, "baseColor", d.material.base_color, e) && getVec3(*m, "emissive", d.material.emissive, e);
		getFloat(*m, "roughness", d.material.roughness);
		getFloat(*m, "metalness", d.material.metalness);
		getString(*m, "texture", tex);
	}
	for (int k = 0; k < 3; ++k) d.rotation[k] = radiansf(rotationDeg[k]);
	return ok;
}

// ---- the streaming reader for large scene files -----------------------------------------------------------------------
// A million-object scene is a 400 MB file of which all but a few hundred bytes is the "objects" array.  Building a DOM of it
// (tens of millions of nodes) and walking that was most of the load time.  For large files the array is instead (1) located by
// a scan of the top-level object, (2) delimited and split at its top-level commas by a two-pass structural scan that all cores
// share (pass 1: every chunk's bracket-depth change and string state under both hypotheses "starts inside / outside a string";
// a serial prefix picks the true one; pass 2: the commas at depth 0), (3) read element by element, in parallel, by a scanner that
// writes the description directly (fastObject).  The scanner accepts exactly the plain form a scene file has - known keys, each
// at most once, strings without escapes, numbers, three-element arrays - and returns false on ANYTHING else; such an element goes
// through the DOM (parseJsonSpan + convertObject), so quirks, messages and error texts stay those of the reference loader.  The
// rest of the document (camera, skybox, whatever else) is parsed as a DOM from the text with the array cut out.  If any step
// is unsure, the whole file takes the DOM path.
struct Cursor
{
	const char *s;
	size_t i, e;
	void ws() { while (i < e && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i; }
	bool eat(char c) { ws(); if (i < e && s[i] == c) { ++i; return true; } return false; }
	// a string without escapes: [b, b + len) is its content
	bool plainString(const char *&b, size_t &len)
	{
		ws();
		if (i >= e || s[i] != '"') return false;
		const size_t start = ++i;
		while (i < e && s[i] != '"')
		{
			if (s[i] == '\\' || (unsigned char)s[i] < 0x20) return false;
			++i;
		}
		if (i >= e) return false;
		b = s + start;
		len = i - start;
		++i;
		return true;
	}
	bool number(float &out, bool *isFloatOut = nullptr)
	{
		ws();
		if (i >= e || !(s[i] == '-' || (s[i] >= '0' && s[i] <= '9'))) return false;
		double v;
		bool isFloat;
		if (!scanJsonNumber(s, e, i, v, isFloat)) return false;
		out = float(v);
		if (isFloatOut) *isFloatOut = isFloat;
		return true;
	}
	bool vec3(float out[3])
	{
		if (!eat('[')) return false;
		for (int k = 0; k < 3; ++k)
		{
			if (k && !eat(',')) return false;
			if (!number(out[k])) return false;
		}
		return eat(']');
	}
};
inline bool keyIs(const char *b, size_t len, const char *lit) { return strlen(lit) == len && memcmp(b, lit, len) == 0; }

bool fastObject(const char *s, size_t begin, size_t end, pt_object_desc &d, std::string &tex)
{
	Cursor c{ s, begin, end };
	objectDefaults(d);
	float rotationDeg[3] = { 0.0f, 0.0f, 0.0f };
	unsigned seen = 0;
	if (!c.eat('{')) return false;
	if (!c.eat('}'))
	{
		while (true)
		{
			const char *k;
			size_t kl;
			if (!c.plainString(k, kl) || !c.eat(':')) return false;
			unsigned bit;
			if (keyIs(k, kl, "name")) { bit = 1; const char *v; size_t vl; if (!c.plainString(v, vl)) return false; }
			else if (keyIs(k, kl, "type"))
			{
				bit = 2;
				const char *v; size_t vl;
				if (!c.plainString(v, vl)) return false;
				bool found = false;
				for (uint32_t t = 0; t < 7; ++t) if (keyIs(v, vl, kShapeNames[t])) { d.type = t; found = true; }
				if (!found) return false; // the DOM path prints the reference's message
			}
			else if (keyIs(k, kl, "position")) { bit = 4; if (!c.vec3(d.position)) return false; }
			else if (keyIs(k, kl, "rotation")) { bit = 8; if (!c.vec3(rotationDeg)) return false; }
			else if (keyIs(k, kl, "scale")) { bit = 16; if (!c.vec3(d.scale)) return false; }
			else if (keyIs(k, kl, "material"))
			{
				bit = 32;
				unsigned mseen = 0;
				if (!c.eat('{')) return false;
				if (!c.eat('}'))
					while (true)
					{
						const char *mk;
						size_t ml;
						if (!c.plainString(mk, ml) || !c.eat(':')) return false;
						unsigned mbit;
						if (keyIs(mk, ml, "type"))
						{
							mbit = 1;
							const char *v; size_t vl;
							if (!c.plainString(v, vl)) return false;
							if (keyIs(v, vl, "LAMBERT")) d.material.type = PT_LAMBERT;
							else if (keyIs(v, vl, "GGX")) d.material.type = PT_GGX;
							else if (keyIs(v, vl, "LAMBERT_GGX")) d.material.type = PT_LAMBERT_GGX;
							else return false;
						}
						else if (keyIs(mk, ml, "baseColor")) { mbit = 2; if (!c.vec3(d.material.base_color)) return false; }
						else if (keyIs(mk, ml, "emissive")) { mbit = 4; if (!c.vec3(d.material.emissive)) return false; }
						else if (keyIs(mk, ml, "roughness") || keyIs(mk, ml, "metalness"))
						{
							const bool rough = mk[0] == 'r';
							mbit = rough ? 8 : 16;
							float v;
							bool isFloat;
							if (!c.number(v, &isFloat)) return false;
							if (isFloat) (rough ? d.material.roughness : d.material.metalness) = v; // integer literals are ignored (Q5)
						}
						else if (keyIs(mk, ml, "texture")) { mbit = 32; const char *v; size_t vl; if (!c.plainString(v, vl)) return false; tex.assign(v, vl); }
						else return false;
						if (mseen & mbit) return false; // a repeated key: the DOM knows which one wins
						mseen |= mbit;
						if (c.eat(',')) continue;
						if (c.eat('}')) break;
						return false;
					}
			}
			else return false;
			if (seen & bit) return false;
			seen |= bit;
			if (c.eat(',')) continue;
			if (c.eat('}')) break;
			return false;
		}
	}
	c.ws();
	if (c.i != end) return false;
	for (int k = 0; k < 3; ++k) d.rotation[k] = radiansf(rotationDeg[k]);
	return true;
}

// structural scan of text[begin, end): bracket depth and string state; `onComma(pos)` for commas at relative depth `wantDepth`
// (pass 2), stops behind the ']' / '}' that takes the depth below `floorDepth` when stopAtFloor.  Returns the position it stopped at.
struct ScanState { bool inString; long depth; long minDepth; };
template <typename F>
size_t structuralScan(const char *s, size_t begin, size_t end, ScanState &st, long commaDepth, bool stopAtFloor, long floorDepth, F onComma)
{
	static const struct Structural { bool is[256]; Structural() : is() { for (const char *q = "\"[]{},\\"; *q; ++q) is[(unsigned char)*q] = true; } } structural;
	size_t j = begin;
	while (j < end)
	{
		if (st.inString)
		{
			while (j < end && s[j] != '"' && s[j] != '\\') ++j;
			if (j >= end) break;
			if (s[j] == '\\') { j += 2; continue; } // (chunk boundaries never fall behind a backslash)
			st.inString = false;
			++j;
			continue;
		}
		while (j < end && !structural.is[(unsigned char)s[j]]) ++j;
		if (j >= end) break;
		const char ch = s[j];
		if (ch == '"') st.inString = true;
		else if (ch == '[' || ch == '{') ++st.depth;
		else if (ch == ']' || ch == '}')
		{
			--st.depth;
			if (st.depth < st.minDepth) st.minDepth = st.depth;
			if (stopAtFloor && st.depth < floorDepth) return j;
		}
		else if (ch == ',' && st.depth == commaDepth) onComma(j);
		++j;
	}
	return end;
}

// Locates "objects": [ ... ] in the top-level object and splits it.  arrBegin / arrEnd = positions of its brackets.
bool splitObjectsArray(const char *s, const size_t n, size_t &arrBegin, size_t &arrEnd, std::vector<size_t> &starts)
{
	// (1) the key, by a scan of the top-level object's members
	size_t i = 0;
	if (n >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) i = 3;
	Cursor c{ s, i, n };
	if (!c.eat('{')) return false;
	bool found = false;
	while (!found)
	{
		const char *k;
		size_t kl;
		if (!c.plainString(k, kl) || !c.eat(':')) return false;
		c.ws();
		if (keyIs(k, kl, "objects"))
		{
			if (c.i >= n || s[c.i] != '[') return false;
			arrBegin = c.i;
			found = true;
			break;
		}
		// skip this member's value (small: camera, skybox, ...)
		ScanState st{ false, 0, 0 };
		if (c.i < n && (s[c.i] == '{' || s[c.i] == '['))
		{
			// a nested value (the camera): up to the bracket that closes it; one that does not close within a megabyte is not a
			// scene file as we know it - the DOM handles it
			const size_t limit = std::min(n, c.i + (size_t(1) << 20));
			const size_t stop = structuralScan(s, c.i, limit, st, LONG_MIN, true, 1, [](size_t) {});
			if (stop >= limit) return false;
			c.i = stop + 1;
		}
		else if (c.i < n && s[c.i] == '"') { const char *v; size_t vl; if (!c.plainString(v, vl)) return false; }
		else
		{
			while (c.i < n && s[c.i] != ',' && s[c.i] != '}') ++c.i; // number / true / false / null
		}
		if (!c.eat(',')) return false;
	}
	// (2) whether a chunk starts inside a string: the parity of the unescaped quotes in front of it (memchr speed; a chunk never
	// starts behind a backslash, so a run of backslashes is never cut)
	const size_t from = arrBegin + 1;
	int threads = 1;
#ifdef _OPENMP
	threads = omp_get_max_threads();
#endif
	const size_t nChunks = std::max<size_t>(1, std::min<size_t>(size_t(threads) * 8, (n - from) >> 16));
	std::vector<size_t> bound(nChunks + 1);
	for (size_t k = 0; k <= nChunks; ++k)
	{
		size_t b = from + (n - from) * k / nChunks;
		while (b > from && b < n && s[b - 1] == '\\') ++b; // never start a chunk on an escaped character
		bound[k] = std::min(b, n);
	}
	bound[nChunks] = n;
	std::vector<size_t> quotes(nChunks);
#pragma omp parallel for schedule(dynamic, 1)
	for (long k = 0; k < long(nChunks); ++k)
	{
		size_t count = 0;
		const char *p = s + bound[k], *e = s + bound[k + 1];
		while (p < e && (p = static_cast<const char *>(memchr(p, '"', size_t(e - p)))) != nullptr)
		{
			size_t slashes = 0;
			for (const char *q = p; q > s + bound[k] && q[-1] == '\\'; --q) ++slashes;
			if ((slashes & 1u) == 0u) ++count;
			++p;
		}
		quotes[size_t(k)] = count;
	}
	// (3) ONE structural scan per chunk, from depth 0: the commas at the lowest depth the chunk reaches are its candidates for the
	// array's top-level commas (they are the ones exactly when that lowest depth is the array's own level, which the prefix over
	// the chunks' depth changes tells afterwards)
	struct ChunkScan { long depth, minDepth; std::vector<size_t> commas; };
	std::vector<ChunkScan> scan(nChunks);
	{
		std::vector<char> startsInString(nChunks);
		size_t q = 0;
		for (size_t k = 0; k < nChunks; ++k) { startsInString[k] = char(q & 1u); q += quotes[k]; }
#pragma omp parallel for schedule(dynamic, 1)
		for (long k = 0; k < long(nChunks); ++k)
		{
			ChunkScan &cs = scan[size_t(k)];
			bool inString = startsInString[size_t(k)] != 0;
			long depth = 0, minDepth = 0;
			static const struct Structural { bool is[256]; Structural() : is() { for (const char *z = "\"[]{},\\"; *z; ++z) is[(unsigned char)*z] = true; } } structural;
			size_t j = bound[k];
			const size_t end = bound[k + 1];
			while (j < end)
			{
				if (inString)
				{
					while (j < end && s[j] != '"' && s[j] != '\\') ++j;
					if (j >= end) break;
					if (s[j] == '\\') { j += 2; continue; }
					inString = false;
					++j;
					continue;
				}
				while (j < end && !structural.is[(unsigned char)s[j]]) ++j;
				if (j >= end) break;
				const char ch = s[j];
				if (ch == '"') inString = true;
				else if (ch == '[' || ch == '{') ++depth;
				else if (ch == ']' || ch == '}') { if (--depth < minDepth) { minDepth = depth; cs.commas.clear(); } }
				else if (ch == ',' && depth == minDepth) cs.commas.push_back(j);
				++j;
			}
			cs.depth = depth;
			cs.minDepth = minDepth;
		}
	}
	// serial prefix: absolute depth at the start of every chunk, and the chunk in which the array closes (depth -1)
	std::vector<long> depthAt(nChunks);
	long depth = 0;
	size_t lastChunk = nChunks;
	for (size_t k = 0; k < nChunks; ++k)
	{
		depthAt[k] = depth;
		if (depth + scan[k].minDepth < 0) { lastChunk = k; break; }
		depth += scan[k].depth;
	}
	if (lastChunk == nChunks) return false; // the array never closes
	// the closing chunk holds text behind the array as well: scanned once more, with its known start state, up to the bracket
	std::vector<size_t> lastCommas;
	size_t closeAt = 0;
	{
		size_t q = 0;
		for (size_t k = 0; k < lastChunk; ++k) q += quotes[k];
		ScanState st{ (q & 1u) != 0u, depthAt[lastChunk], depthAt[lastChunk] };
		closeAt = structuralScan(s, bound[lastChunk], bound[lastChunk + 1], st, 0, true, 0, [&lastCommas](size_t pos) { lastCommas.push_back(pos); });
	}
	if (closeAt >= n || s[closeAt] != ']') return false;
	arrEnd = closeAt;
	starts.clear();
	starts.push_back(from);
	for (size_t k = 0; k < lastChunk; ++k)
		if (depthAt[k] + scan[k].minDepth == 0) // the chunk reaches the array's own level
			for (size_t pos : scan[k].commas) starts.push_back(pos + 1);
	for (size_t pos : lastCommas) starts.push_back(pos + 1);
	return true;
}
} // namespace

bool parseSceneText(const std::string &text, float aspect, ParsedScene &out, std::string &err) { return parseSceneText(text.data(), text.size(), aspect, out, err); }
bool parseSceneText(const char *textData, const size_t textSize, float aspect, ParsedScene &out, std::string &err)
{
	out = ParsedScene();
	auto textureHandle = [&](const std::string &p) -> uint32_t
	{
		if (p.empty()) return 0;
		for (size_t i = 0; i < out.texturePaths.size(); ++i)
			if (out.texturePaths[i] == p) return uint32_t(i + 1);
		out.texturePaths.push_back(p);
		return uint32_t(out.texturePaths.size());
	};
	// per object, by all cores: everything except the two things that depend on the order of the objects (texture handles
	// are given out in first-use order, the loader's messages are printed in object order) - those follow in one ordered pass
	std::vector<std::string> texOf;
	std::vector<std::vector<std::string>> msgOf;
	long firstBad = -1;
	std::string firstErr;
	auto orderedPass = [&](long count) -> bool
	{
		const long stopAt = firstBad < 0 ? count : firstBad + 1;
		for (long i = 0; i < stopAt; ++i)
		{
			for (std::string &msg : msgOf[size_t(i)]) out.messages.push_back(std::move(msg));
			if (i == firstBad) { err = firstErr; return false; } // a type error ends the load at the first bad object, messages printed up to there
			if (!texOf[size_t(i)].empty()) out.objects[size_t(i)].material.texture = textureHandle(texOf[size_t(i)]);
		}
		return true;
	};

	JsonValue root;
	bool streamed = false;
	size_t arrBegin = 0, arrEnd = 0;
	std::vector<size_t> starts;
	LoaderLap lap;
	if (textSize >= (size_t(4) << 20) && splitObjectsArray(textData, textSize, arrBegin, arrEnd, starts))
	{
		lap("split array");
		// the rest of the document, with the array cut out
		std::string rest;
		rest.reserve(arrBegin + (textSize - arrEnd) + 2);
		rest.append(textData, arrBegin);
		rest += "[]";
		rest.append(textData + arrEnd + 1, textSize - arrEnd - 1);
		std::string restErr;
		const JsonValue *ro = nullptr;
		if (parseJson(rest, root, restErr) && (ro = root.find("objects")) != nullptr && ro->kind == JsonValue::Array && ro->arr().empty())
		{
			const char *s = textData;
			// "[ ]": one blank element
			size_t count = starts.size();
			if (count == 1)
			{
				size_t b = starts[0];
				while (b < arrEnd && (s[b] == ' ' || s[b] == '\t' || s[b] == '\n' || s[b] == '\r')) ++b;
				if (b == arrEnd) count = 0;
			}
			out.hasObjectsArray = true;
			out.objects.resize(count);
			texOf.assign(count, std::string());
			msgOf.assign(count, std::vector<std::string>());
			bool syntaxError = false;
			long bad = long(count);
#pragma omp parallel for schedule(dynamic, 1024)
			for (long i = 0; i < long(count); ++i)
			{
				if (syntaxError) continue;
				const size_t b = starts[size_t(i)], e = size_t(i) + 1 < count ? starts[size_t(i) + 1] - 1 : arrEnd;
				if (fastObject(s, b, e, out.objects[size_t(i)], texOf[size_t(i)])) continue;
				// anything out of the ordinary: this element through the DOM, like the small-file path
				texOf[size_t(i)].clear();
				JsonValue v;
				std::string e2;
				if (!parseJsonSpan(s, b, e, v)) { syntaxError = true; continue; }
				if (!convertObject(v, out.objects[size_t(i)], texOf[size_t(i)], msgOf[size_t(i)], e2))
				{
#pragma omp critical(ptb_scene_loader_error)
					if (i < bad) { bad = i; firstErr = e2; }
				}
			}
			lap("objects");
			if (!syntaxError)
			{
				streamed = true;
				firstBad = bad < long(count) ? bad : -1;
				if (!orderedPass(long(count))) return false;
			}
		}
		if (!streamed)
		{
			// something was not as expected: start over with the DOM of the whole text (its errors carry the right byte offsets)
			out = ParsedScene();
			root = JsonValue();
		}
	}
	if (!streamed)
	{
		if (!parseJson(textData, textSize, root, err)) return false;
		const JsonValue *objs = root.find("objects");
		if (objs && objs->kind == JsonValue::Array)
		{
			out.hasObjectsArray = true;
			const std::vector<JsonValue> &list = objs->arr();
			const long count = long(list.size());
			out.objects.resize(size_t(count));
			texOf.assign(size_t(count), std::string());
			msgOf.assign(size_t(count), std::vector<std::string>());
			long bad = count;
#pragma omp parallel for schedule(static) if (count > 4096)
			for (long i = 0; i < count; ++i)
			{
				std::string e;
				if (!convertObject(list[size_t(i)], out.objects[size_t(i)], texOf[size_t(i)], msgOf[size_t(i)], e))
				{
#pragma omp critical(ptb_scene_loader_error)
					if (i < bad) { bad = i; firstErr = e; }
				}
			}
			firstBad = bad < count ? bad : -1;
			if (!orderedPass(count)) return false;
			// give the DOM of a large list back in parallel too (tens of millions of small nodes)
			if (count > 4096)
			{
				std::vector<JsonValue> &mut = const_cast<JsonValue *>(objs)->arr();
#pragma omp parallel for schedule(static)
				for (long i = 0; i < count; ++i) mut[size_t(i)] = JsonValue();
			}
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
	f.close();
	// the file is MAPPED, not read: a million-object scene is hundreds of megabytes, and copying it out of the page cache into a
	// zero-filled buffer was a third of the load time; the loader's threads fault the pages in as they scan them
	LoaderLap lapRead;
	const int fd = open(path, O_RDONLY);
	struct stat st;
	if (fd < 0 || fstat(fd, &st) != 0)
	{
		if (fd >= 0) close(fd);
		err = std::string("Failed to open input file: ") + path;
		if (errCode) *errCode = PT_E_IO;
		return false;
	}
	const size_t size = size_t(st.st_size);
	void *map = size ? mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
	std::string small;
	const char *text = "";
	if (size && map == MAP_FAILED)
	{
		// (a file system that cannot map: read it)
		map = nullptr;
		small.resize(size);
		size_t got = 0;
		while (got < size) { const ssize_t r = pread(fd, &small[got], size - got, off_t(got)); if (r <= 0) break; got += size_t(r); }
		small.resize(got);
		text = small.data();
	}
	else if (size)
	{
		madvise(map, size, MADV_SEQUENTIAL | MADV_WILLNEED);
		text = static_cast<const char *>(map);
	}
	close(fd);
	lapRead("map file");
	const size_t textSize = map ? size : small.size();
	const bool ok = parseSceneText(text, textSize, aspect, out, err);
	if (map) munmap(map, size);
	if (!ok)
	{
		if (errCode) *errCode = PT_E_PARSE;
		return false;
	}
	return true;
}
} // namespace ptb
