// json_min.h — a small JSON DOM reader, enough for the scene schema.  It keeps the one distinction
// the reference's loader depends on: whether a number was written as a float literal ('.', 'e' or 'E' present) or an
// integer literal (nlohmann's is_number_float(), used by reference SceneLoader.cpp:163-171 — quirk Q5).
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace ptb
{
struct JsonValue
{
	enum Kind { Null, Bool, Int, Float, String, Array, Object } kind = Null;
	bool b = false;
	double num = 0.0;
	std::string str;
	std::vector<JsonValue> arr;
	std::vector<std::pair<std::string, JsonValue>> obj; // insertion order; a repeated key replaces the earlier value (as nlohmann)

	bool isNumber() const { return kind == Int || kind == Float; }
	const JsonValue *find(const char *key) const
	{
		if (kind != Object) return nullptr;
		for (const auto &kv : obj)
			if (kv.first == key) return &kv.second;
		return nullptr;
	}
};

// Parses `text`; on failure returns false and describes the error (with byte offset) in `err`.
bool parseJson(const std::string &text, JsonValue &out, std::string &err);
} // namespace ptb
