// json.hpp — minimal JSON value + parser/serialiser for the recconf subset the host mirror reads.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace pairec {

class Json {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };
  Type type = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

  bool is_null() const { return type == Null; }
  const Json* find(const std::string& k) const {
    if (type != Object) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Json& operator[](const std::string& k) const {
    static const Json nul;
    const Json* p = find(k);
    return p ? *p : nul;
  }
  std::string as_string(const std::string& d = "") const { return type == String ? str : d; }
  double as_number(double d = 0) const { return type == Number ? num : d; }
  int as_int(int d = 0) const { return type == Number ? (int)num : d; }
  bool as_bool(bool d = false) const { return type == Bool ? b : d; }

  static bool parse(const std::string& s, Json* out, std::string* err) {
    size_t i = 0;
    if (!parse_value(s, i, out, err)) return false;
    skip(s, i);
    if (i != s.size()) { if (err) *err = "trailing characters at " + std::to_string(i); return false; }
    return true;
  }
  static std::string quote(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
      if (c == '"') o += "\\\"";
      else if (c == '\\') o += "\\\\";
      else if (c == '\n') o += "\\n";
      else if (c == '\t') o += "\\t";
      else if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); o += buf; }
      else o += (char)c;
    }
    return o + "\"";
  }
  static std::string number(double v) {
    char buf[40];
    snprintf(buf, sizeof buf, "%.17g", v);
    return buf;
  }

 private:
  static void skip(const std::string& s, size_t& i) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i;
  }
  static bool fail(std::string* err, const std::string& m, size_t i) {
    if (err) *err = m + " at offset " + std::to_string(i);
    return false;
  }
  static bool parse_string(const std::string& s, size_t& i, std::string* out, std::string* err) {
    if (s[i] != '"') return fail(err, "expected string", i);
    ++i;
    out->clear();
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\') {
        ++i;
        if (i >= s.size()) return fail(err, "bad escape", i);
        switch (s[i]) {
          case 'n': *out += '\n'; break;
          case 't': *out += '\t'; break;
          case 'r': *out += '\r'; break;
          case 'b': *out += '\b'; break;
          case 'f': *out += '\f'; break;
          case 'u': {
            if (i + 4 >= s.size()) return fail(err, "bad \\u escape", i);
            unsigned cp = (unsigned)strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16);
            i += 4;
            if (cp < 0x80) *out += (char)cp;
            else if (cp < 0x800) { *out += (char)(0xC0 | (cp >> 6)); *out += (char)(0x80 | (cp & 0x3F)); }
            else { *out += (char)(0xE0 | (cp >> 12)); *out += (char)(0x80 | ((cp >> 6) & 0x3F)); *out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: *out += s[i];
        }
        ++i;
      } else {
        *out += s[i++];
      }
    }
    if (i >= s.size()) return fail(err, "unterminated string", i);
    ++i;
    return true;
  }
  static bool parse_value(const std::string& s, size_t& i, Json* out, std::string* err) {
    skip(s, i);
    if (i >= s.size()) return fail(err, "unexpected end", i);
    const char c = s[i];
    if (c == '{') {
      out->type = Object;
      ++i;
      skip(s, i);
      if (i < s.size() && s[i] == '}') { ++i; return true; }
      for (;;) {
        skip(s, i);
        std::string k;
        if (i >= s.size() || !parse_string(s, i, &k, err)) return false;
        skip(s, i);
        if (i >= s.size() || s[i] != ':') return fail(err, "expected ':'", i);
        ++i;
        Json v;
        if (!parse_value(s, i, &v, err)) return false;
        out->obj.emplace_back(std::move(k), std::move(v));
        skip(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == '}') { ++i; return true; }
        return fail(err, "expected ',' or '}'", i);
      }
    }
    if (c == '[') {
      out->type = Array;
      ++i;
      skip(s, i);
      if (i < s.size() && s[i] == ']') { ++i; return true; }
      for (;;) {
        Json v;
        if (!parse_value(s, i, &v, err)) return false;
        out->arr.push_back(std::move(v));
        skip(s, i);
        if (i < s.size() && s[i] == ',') { ++i; continue; }
        if (i < s.size() && s[i] == ']') { ++i; return true; }
        return fail(err, "expected ',' or ']'", i);
      }
    }
    if (c == '"') { out->type = String; return parse_string(s, i, &out->str, err); }
    if (s.compare(i, 4, "true") == 0) { out->type = Bool; out->b = true; i += 4; return true; }
    if (s.compare(i, 5, "false") == 0) { out->type = Bool; out->b = false; i += 5; return true; }
    if (s.compare(i, 4, "null") == 0) { out->type = Null; i += 4; return true; }
    char* end = nullptr;
    const double v = strtod(s.c_str() + i, &end);
    if (end == s.c_str() + i) return fail(err, "unexpected character", i);
    const std::string tok(s.c_str() + i, (size_t)(end - (s.c_str() + i)));
    out->type = Number;
    out->num = v;
    out->is_int = tok.find_first_of(".eE") == std::string::npos;
    i = (size_t)(end - s.c_str());
    return true;
  }
};

}  // namespace pairec
