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
	enum Kind : unsigned char { Null, Bool, Int, Float, String, Array, Object } kind = Null;
	bool b = false;
	double num = 0.0;
	// strings, arrays and objects keep their payload behind one pointer: a number - most of the values of a scene file,
	// tens of millions in a million-object scene - is 24 bytes and no allocation
	struct Payload;
	std::unique_ptr<Payload> payload;

	JsonValue();
	~JsonValue();
	JsonValue(JsonValue &&) noexcept;
	JsonValue &operator=(JsonValue &&) noexcept;
	JsonValue(const JsonValue &) = delete;
	JsonValue &operator=(const JsonValue &) = delete;

	std::string &str();
	std::vector<JsonValue> &arr();
	std::vector<std::pair<std::string, JsonValue>> &obj(); // insertion order; a repeated key replaces the earlier value (as nlohmann)
	const std::string &str() const;
	const std::vector<JsonValue> &arr() const;
	const std::vector<std::pair<std::string, JsonValue>> &obj() const;

	bool isNumber() const { return kind == Int || kind == Float; }
	const JsonValue *find(const char *key) const;
};
struct JsonValue::Payload
{
	std::string str;
	std::vector<JsonValue> arr;
	std::vector<std::pair<std::string, JsonValue>> obj;
};
inline JsonValue::JsonValue() = default;
inline JsonValue::~JsonValue() = default;
inline JsonValue::JsonValue(JsonValue &&) noexcept = default;
inline JsonValue &JsonValue::operator=(JsonValue &&) noexcept = default;
inline std::string &JsonValue::str() { if (!payload) payload.reset(new Payload()); return payload->str; }
inline std::vector<JsonValue> &JsonValue::arr() { if (!payload) payload.reset(new Payload()); return payload->arr; }
inline std::vector<std::pair<std::string, JsonValue>> &JsonValue::obj() { if (!payload) payload.reset(new Payload()); return payload->obj; }
inline const std::string &JsonValue::str() const { static const std::string e; return payload ? payload->str : e; }
inline const std::vector<JsonValue> &JsonValue::arr() const { static const std::vector<JsonValue> e; return payload ? payload->arr : e; }
inline const std::vector<std::pair<std::string, JsonValue>> &JsonValue::obj() const { static const std::vector<std::pair<std::string, JsonValue>> e; return payload ? payload->obj : e; }
inline const JsonValue *JsonValue::find(const char *key) const
{
	if (kind != Object) return nullptr;
	for (const auto &kv : obj())
		if (kv.first == key) return &kv.second;
	return nullptr;
}

// Parses `text`; on failure returns false and describes the error (with byte offset) in `err`.
bool parseJson(const std::string &text, JsonValue &out, std::string &err);
bool parseJson(const char *text, size_t n, JsonValue &out, std::string &err); // (the text need not be null-terminated: a mapped file)
// Parses the ONE value that fills text[begin, end) (blanks around it allowed), sequentially.  For the streaming scene reader.
bool parseJsonSpan(const char *text, size_t begin, size_t end, JsonValue &out);
// Scans one JSON number at text[i...] (i < n, text[i] is '-' or a digit): value as a double (as parseJson produces it), whether
// it was written as a float literal (quirk Q5), i moved behind it.  False on a malformed number.
bool scanJsonNumber(const char *text, size_t n, size_t &i, double &value, bool &isFloat);
} // namespace ptb
