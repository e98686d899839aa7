// json_min.hpp -- a small strict JSON reader for the stateline server config.
// The reference reads its config with nlohmann::json (external/, src/app/jsonsettings.hpp:22-35); the
// engine only needs objects, arrays, numbers, strings, booleans and null, so it carries its own reader.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace slgpu_json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<Value> arr;
  std::map<std::string, Value> obj;

  const Value* find(const std::string& key) const {
    if (kind != Object) return nullptr;
    auto it = obj.find(key);
    return it == obj.end() ? nullptr : &it->second;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text) {}
  Value parse() {
    Value v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const std::string& what) const {
    throw std::runtime_error("JSON: " + what + " at offset " + std::to_string(i_));
  }
  void ws() {
    while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_;
  }
  bool eat(char c) {
    ws();
    if (i_ < s_.size() && s_[i_] == c) {
      ++i_;
      return true;
    }
    return false;
  }
  void expect(char c) {
    if (!eat(c)) fail(std::string("expected '") + c + "'");
  }
  bool literal(const char* lit) {
    size_t n = 0;
    while (lit[n]) ++n;
    if (s_.compare(i_, n, lit) == 0) {
      i_ += n;
      return true;
    }
    return false;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (i_ < s_.size() && s_[i_] != '"') {
      char c = s_[i_++];
      if (c == '\\') {
        if (i_ >= s_.size()) fail("bad escape");
        char e = s_[i_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (i_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned code = (unsigned)std::strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
            i_ += 4;
            if (code < 0x80) out += (char)code;
            else if (code < 0x800) { out += (char)(0xC0 | (code >> 6)); out += (char)(0x80 | (code & 0x3F)); }
            else { out += (char)(0xE0 | (code >> 12)); out += (char)(0x80 | ((code >> 6) & 0x3F)); out += (char)(0x80 | (code & 0x3F)); }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return out;
  }
  Value value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    Value v;
    const char c = s_[i_];
    if (c == '{') {
      ++i_;
      v.kind = Value::Object;
      if (eat('}')) return v;
      do {
        ws();
        std::string k = string();
        expect(':');
        v.obj[k] = value();
      } while (eat(','));
      expect('}');
    } else if (c == '[') {
      ++i_;
      v.kind = Value::Array;
      if (eat(']')) return v;
      do v.arr.push_back(value());
      while (eat(','));
      expect(']');
    } else if (c == '"') {
      v.kind = Value::String;
      v.str = string();
    } else if (literal("true")) {
      v.kind = Value::Bool;
      v.b = true;
    } else if (literal("false")) {
      v.kind = Value::Bool;
      v.b = false;
    } else if (literal("null")) {
      v.kind = Value::Null;
    } else {
      const char* start = s_.c_str() + i_;
      char* end = nullptr;
      v.num = std::strtod(start, &end);
      if (end == start) fail("unexpected character");
      i_ += (size_t)(end - start);
      v.kind = Value::Number;
    }
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace slgpu_json
