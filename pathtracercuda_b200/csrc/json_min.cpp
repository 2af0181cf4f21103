#include "json_min.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ptb
{
namespace
{
struct Parser
{
	const char *s;
	size_t n, i = 0;
	std::string err;
	int depth = 0;

	bool fail(const char *m)
	{
		if (err.empty()) err = std::string("JSON parse error at byte ") + std::to_string(i) + ": " + m;
		return false;
	}
	void ws() { while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i; }

	static void appendUtf8(std::string &o, unsigned cp)
	{
		if (cp < 0x80) o += char(cp);
		else if (cp < 0x800) { o += char(0xC0 | (cp >> 6)); o += char(0x80 | (cp & 0x3F)); }
		else if (cp < 0x10000) { o += char(0xE0 | (cp >> 12)); o += char(0x80 | ((cp >> 6) & 0x3F)); o += char(0x80 | (cp & 0x3F)); }
		else { o += char(0xF0 | (cp >> 18)); o += char(0x80 | ((cp >> 12) & 0x3F)); o += char(0x80 | ((cp >> 6) & 0x3F)); o += char(0x80 | (cp & 0x3F)); }
	}
	bool hex4(unsigned &v)
	{
		if (i + 4 > n) return fail("truncated \\u escape");
		v = 0;
		for (int k = 0; k < 4; ++k)
		{
			char c = s[i++];
			v <<= 4;
			if (c >= '0' && c <= '9') v |= unsigned(c - '0');
			else if (c >= 'a' && c <= 'f') v |= unsigned(c - 'a' + 10);
			else if (c >= 'A' && c <= 'F') v |= unsigned(c - 'A' + 10);
			else return fail("bad \\u escape");
		}
		return true;
	}
	bool string(std::string &o)
	{
		++i; // opening quote
		while (true)
		{
			if (i >= n) return fail("unterminated string");
			unsigned char c = (unsigned char)s[i++];
			if (c == '"') return true;
			if (c < 0x20) return fail("control character in string");
			if (c != '\\') { o += char(c); continue; }
			if (i >= n) return fail("unterminated escape");
			char e = s[i++];
			switch (e)
			{
			case '"': o += '"'; break;
			case '\\': o += '\\'; break;
			case '/': o += '/'; break;
			case 'b': o += '\b'; break;
			case 'f': o += '\f'; break;
			case 'n': o += '\n'; break;
			case 'r': o += '\r'; break;
			case 't': o += '\t'; break;
			case 'u':
			{
				unsigned cp = 0;
				if (!hex4(cp)) return false;
				if (cp >= 0xD800 && cp <= 0xDBFF && i + 1 < n && s[i] == '\\' && s[i + 1] == 'u')
				{
					i += 2;
					unsigned lo = 0;
					if (!hex4(lo)) return false;
					if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
					else return fail("bad surrogate pair");
				}
				appendUtf8(o, cp);
				break;
			}
			default: return fail("bad escape");
			}
		}
	}
	bool number(JsonValue &v)
	{
		size_t b = i;
		bool isFloat = false;
		if (i < n && s[i] == '-') ++i;
		if (i >= n || s[i] < '0' || s[i] > '9') return fail("bad number");
		if (s[i] == '0') ++i;
		else while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
		if (i < n && s[i] == '.')
		{
			isFloat = true;
			++i;
			if (i >= n || s[i] < '0' || s[i] > '9') return fail("bad fraction");
			while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
		}
		if (i < n && (s[i] == 'e' || s[i] == 'E'))
		{
			isFloat = true;
			++i;
			if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
			if (i >= n || s[i] < '0' || s[i] > '9') return fail("bad exponent");
			while (i < n && s[i] >= '0' && s[i] <= '9') ++i;
		}
		v.kind = isFloat ? JsonValue::Float : JsonValue::Int;
		// Clinger's fast path: a decimal with <= 15 significant digits and a power of ten up to 10^22 is ONE correctly
		// rounded double operation (both operands exact); everything else goes to strtod
		{
			size_t k = b;
			const bool neg = s[k] == '-';
			if (neg) ++k;
			unsigned long long mant = 0;
			int digits = 0, exp10 = 0;
			bool fast = true;
			for (; k < i && s[k] >= '0' && s[k] <= '9'; ++k) { mant = mant * 10 + unsigned(s[k] - '0'); if (mant) ++digits; }
			if (k < i && s[k] == '.')
				for (++k; k < i && s[k] >= '0' && s[k] <= '9'; ++k) { mant = mant * 10 + unsigned(s[k] - '0'); if (mant) ++digits; --exp10; }
			const bool plainDecimal = !(k < i && (s[k] == 'e' || s[k] == 'E'));
			if (digits > 15)
			{
				fast = false;
				// 16 - 19 significant digits without an exponent (what printing a double with 17 digits gives - every number
				// of a generated scene file): the value is the exact rational mant / 10^frac, rounded to double ONCE by a 128-bit
				// integer division - no strtod (which was most of the loader's time on a million-object scene)
				if (plainDecimal && digits <= 19 && exp10 >= -19)
				{
					static const unsigned long long p10u[20] = { 1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
						10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull,
						100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull };
					const unsigned long long den = p10u[-exp10];
					const unsigned __int128 num = (unsigned __int128)mant << 64;
					const unsigned __int128 q = num / den, r = num % den;
					const unsigned long long qh = (unsigned long long)(q >> 64), ql = (unsigned long long)q;
					const int hb = qh ? 127 - __builtin_clzll(qh) : (ql ? 63 - __builtin_clzll(ql) : -1);
					if (hb >= 54)
					{
						int shift = hb - 52;
						unsigned long long m = (unsigned long long)(q >> shift);
						const unsigned __int128 rem = q & ((((unsigned __int128)1) << shift) - 1), half = ((unsigned __int128)1) << (shift - 1);
						if (rem > half || (rem == half && (r != 0 || (m & 1ull)))) { if (++m == (1ull << 53)) { m >>= 1; ++shift; } }
						const double d = ldexp(double(m), shift - 64);
						v.num = neg ? -d : d;
						return true;
					}
				}
			}
			if (fast && k < i && (s[k] == 'e' || s[k] == 'E'))
			{
				++k;
				bool eneg = false;
				if (s[k] == '+' || s[k] == '-') { eneg = s[k] == '-'; ++k; }
				int e = 0;
				for (; k < i && e < 10000; ++k) e = e * 10 + (s[k] - '0');
				exp10 += eneg ? -e : e;
			}
			static const double p10[] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
			if (fast && exp10 >= -22 && exp10 <= 22)
			{
				double d = double(mant);
				d = exp10 < 0 ? d / p10[-exp10] : d * p10[exp10];
				v.num = neg ? -d : d;
				return true;
			}
		}
		char buf[64];
		const size_t len = i - b;
		if (len < sizeof buf)
		{
			memcpy(buf, s + b, len);
			buf[len] = 0;
			v.num = strtod(buf, nullptr);
		}
		else
		{
			std::string tok(s + b, len);
			v.num = strtod(tok.c_str(), nullptr);
		}
		return true;
	}
	bool value(JsonValue &v);

	// A LARGE array (the "objects" list of a million-object scene is hundreds of megabytes) is split at its top-level
	// commas by one cheap structural scan and its elements are parsed by all cores, each with its own Parser over the
	// same text.  Returns false (nothing consumed) when the array is small, when already inside a parallel region, or
	// when anything looks wrong - the sequential path then parses it and reports errors with the usual byte offsets.
	bool parallelArray(JsonValue &v)
	{
#ifdef _OPENMP
		if (n - i < (size_t(1) << 22) || omp_in_parallel() || omp_get_max_threads() < 2) return false;
		std::vector<size_t> starts;
		size_t j = i + 1, end = 0;
		int d = 0;
		starts.push_back(j);
		static const struct Structural { bool is[256]; Structural() : is() { for (const char *q = "\"[]{},"; *q; ++q) is[(unsigned char)*q] = true; } } structural;
		for (; j < n; ++j)
		{
			while (j < n && !structural.is[(unsigned char)s[j]]) ++j; // digits, letters, blanks: nothing to track
			if (j >= n) break;
			const char ch = s[j];
			if (ch == '"')
			{
				for (++j; j < n && s[j] != '"'; ++j)
					if (s[j] == '\\') ++j;
				if (j >= n) return false;
			}
			else if (ch == '[' || ch == '{') ++d;
			else if (ch == ']' || ch == '}')
			{
				if (d == 0) { if (ch != ']') return false; end = j; break; }
				--d;
			}
			else if (ch == ',' && d == 0) starts.push_back(j + 1);
		}
		if (end == 0 || starts.size() < 1024) return false;
		const size_t count = starts.size();
		std::vector<JsonValue> items(count);
		bool bad = false;
#pragma omp parallel for schedule(dynamic, 512)
		for (long k = 0; k < long(count); ++k)
		{
			if (bad) continue;
			const size_t stop = size_t(k) + 1 < count ? starts[k + 1] - 1 : end; // the comma / closing bracket behind the element
			Parser sub{ s, stop };
			sub.i = starts[k];
			sub.depth = depth;
			if (!sub.value(items[k])) { bad = true; continue; }
			sub.ws();
			if (sub.i != stop) bad = true;
		}
		if (bad) return false;
		v.kind = JsonValue::Array;
		v.arr() = std::move(items);
		i = end + 1;
		--depth; // value() counted this level on entry
		return true;
#else
		(void)v;
		return false;
#endif
	}
};

bool Parser::value(JsonValue &v)
	{
		ws();
		if (i >= n) return fail("unexpected end of input");
		if (++depth > 256) return fail("nesting too deep");
		bool ok = true;
		char c = s[i];
		if (c == '{')
		{
			v.kind = JsonValue::Object;
			auto &members = v.obj();
			members.reserve(8);
			++i;
			ws();
			if (i < n && s[i] == '}') { ++i; }
			else
				while (true)
				{
					ws();
					if (i >= n || s[i] != '"') { ok = fail("expected object key"); break; }
					std::string key;
					if (!string(key)) { ok = false; break; }
					ws();
					if (i >= n || s[i] != ':') { ok = fail("expected ':'"); break; }
					++i;
					JsonValue child;
					if (!value(child)) { ok = false; break; }
					bool replaced = false;
					for (auto &kv : members)
						if (kv.first == key) { kv.second = std::move(child); replaced = true; break; }
					if (!replaced) members.emplace_back(std::move(key), std::move(child));
					ws();
					if (i < n && s[i] == ',') { ++i; continue; }
					if (i < n && s[i] == '}') { ++i; break; }
					ok = fail("expected ',' or '}'");
					break;
				}
		}
		else if (c == '[' && parallelArray(v)) {}
		else if (c == '[')
		{
			v.kind = JsonValue::Array;
			auto &items = v.arr();
			items.reserve(4);
			++i;
			ws();
			if (i < n && s[i] == ']') { ++i; }
			else
				while (true)
				{
					JsonValue child;
					if (!value(child)) { ok = false; break; }
					items.push_back(std::move(child));
					ws();
					if (i < n && s[i] == ',') { ++i; continue; }
					if (i < n && s[i] == ']') { ++i; break; }
					ok = fail("expected ',' or ']'");
					break;
				}
		}
		else if (c == '"') { v.kind = JsonValue::String; ok = string(v.str()); }
		else if (c == '-' || (c >= '0' && c <= '9')) ok = number(v);
		else if (n - i >= 4 && !strncmp(s + i, "true", 4)) { v.kind = JsonValue::Bool; v.b = true; i += 4; }
		else if (n - i >= 5 && !strncmp(s + i, "false", 5)) { v.kind = JsonValue::Bool; v.b = false; i += 5; }
		else if (n - i >= 4 && !strncmp(s + i, "null", 4)) { v.kind = JsonValue::Null; i += 4; }
		else ok = fail("unexpected character");
		--depth;
		return ok;
	}
} // namespace

bool parseJsonSpan(const char *text, size_t begin, size_t end, JsonValue &out)
{
	Parser p{ text, end };
	p.i = begin;
	p.depth = 2; // never large enough for the parallel array path to matter; keeps the nesting limit conservative
	if (!p.value(out)) return false;
	p.ws();
	return p.i == end;
}

bool scanJsonNumber(const char *text, size_t n, size_t &i, double &value, bool &isFloat)
{
	Parser p{ text, n };
	p.i = i;
	JsonValue v;
	if (!p.number(v)) return false;
	i = p.i;
	value = v.num;
	isFloat = v.kind == JsonValue::Float;
	return true;
}

bool parseJson(const std::string &text, JsonValue &out, std::string &err) { return parseJson(text.data(), text.size(), out, err); }
bool parseJson(const char *text, size_t n, JsonValue &out, std::string &err)
{
	Parser p{ text, n };
	// nlohmann skips a UTF-8 byte-order mark
	if (p.n >= 3 && (unsigned char)p.s[0] == 0xEF && (unsigned char)p.s[1] == 0xBB && (unsigned char)p.s[2] == 0xBF) p.i = 3;
	if (!p.value(out)) { err = p.err; return false; }
	p.ws();
	if (p.i != p.n) { p.fail("trailing characters"); err = p.err; return false; }
	return true;
}
} // namespace ptb
