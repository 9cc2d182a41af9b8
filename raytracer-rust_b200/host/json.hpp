// json.hpp — minimal JSON DOM for the host-side scene loader (stands in for serde_json, which the reference
// uses in src/tungsten/parser.rs:248-249).  Objects keep insertion order; numbers are doubles.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pth {

struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  bool is_null() const { return kind == Null; }
  bool is_number() const { return kind == Number; }
  bool is_string() const { return kind == String; }
  bool is_array() const { return kind == Array; }
  bool is_object() const { return kind == Object; }
  const Json *find(const std::string &key) const {
    if (kind != Object) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &s) : s_(s) {}
  Json parse() {
    Json v = value();
    ws();
    if (p_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string &s_;
  size_t p_ = 0;
  [[noreturn]] void fail(const char *m) const {
    throw std::runtime_error(std::string("JSON: ") + m + " at byte " + std::to_string(p_));
  }
  void ws() {
    while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) p_++;
  }
  Json value() {
    ws();
    if (p_ >= s_.size()) fail("unexpected end");
    char c = s_[p_];
    Json v;
    if (c == '{') {
      v.kind = Json::Object;
      p_++;
      ws();
      if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
      for (;;) {
        ws();
        if (p_ >= s_.size() || s_[p_] != '"') fail("expected key");
        std::string k = string();
        ws();
        if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
        p_++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == '}') { p_++; break; }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      v.kind = Json::Array;
      p_++;
      ws();
      if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
        if (p_ < s_.size() && s_[p_] == ']') { p_++; break; }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.kind = Json::String;
      v.str = string();
    } else if (s_.compare(p_, 4, "true") == 0) {
      v.kind = Json::Bool; v.b = true; p_ += 4;
    } else if (s_.compare(p_, 5, "false") == 0) {
      v.kind = Json::Bool; v.b = false; p_ += 5;
    } else if (s_.compare(p_, 4, "null") == 0) {
      p_ += 4;
    } else {
      const char *start = s_.c_str() + p_;
      char *end = nullptr;
      double d = std::strtod(start, &end);
      if (end == start) fail("unexpected character");
      v.kind = Json::Number;
      v.num = d;
      p_ += (size_t)(end - start);
    }
    return v;
  }
  std::string string() {
    std::string out;
    p_++;  // opening quote
    while (p_ < s_.size() && s_[p_] != '"') {
      char c = s_[p_++];
      if (c == '\\') {
        if (p_ >= s_.size()) fail("bad escape");
        char e = s_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (p_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (p_ >= s_.size()) fail("unterminated string");
    p_++;
    return out;
  }
};

}  // namespace pth
