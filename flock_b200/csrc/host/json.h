// json.h -- a small JSON reader for the reference's serde-JSON physical plans
// (flock/src/runtime/context.rs:366-398 marshal/unmarshal; fixtures in flock/src/tests/data/plan/).
// Numbers keep their integer value when they have one (u64 seeds such as 13714699805381954668 do not
// fit int64), objects keep insertion order.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../internal.h"

namespace flock {

struct Json;
using JsonPtr = std::shared_ptr<Json>;

struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  int64_t i64 = 0;
  bool is_int = false;
  std::string str;
  std::vector<JsonPtr> arr;
  std::vector<std::pair<std::string, JsonPtr>> obj;

  bool is_null() const { return kind == Null; }
  bool is_string() const { return kind == String; }
  bool is_object() const { return kind == Object; }
  bool is_array() const { return kind == Array; }
  const Json* get(const char* key) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return kv.second.get();
    return nullptr;
  }
  const Json& at(const char* key) const {
    const Json* j = get(key);
    if (!j) fg::fail(FLOCKGPU_ERR_INVALID, "plan JSON: missing key \"%s\"", key);
    return *j;
  }
  const std::string& as_string(const char* what) const {
    if (kind != String) fg::fail(FLOCKGPU_ERR_INVALID, "plan JSON: %s is not a string", what);
    return str;
  }
  int64_t as_int(const char* what) const {
    if (kind != Number) fg::fail(FLOCKGPU_ERR_INVALID, "plan JSON: %s is not a number", what);
    return is_int ? i64 : int64_t(num);
  }
};

class JsonParser {
 public:
  explicit JsonParser(const char* text) : p_(text) {}
  JsonPtr parse() {
    JsonPtr v = value();
    ws();
    if (*p_) err("trailing characters");
    return v;
  }

 private:
  const char* p_;
  [[noreturn]] void err(const char* what) { fg::fail(FLOCKGPU_ERR_INVALID, "plan JSON: %s near \"%.24s\"", what, p_); }
  void ws() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
  }
  JsonPtr value() {
    ws();
    auto j = std::make_shared<Json>();
    switch (*p_) {
      case '{': {
        j->kind = Json::Object;
        ++p_;
        ws();
        if (*p_ == '}') { ++p_; return j; }
        while (true) {
          ws();
          if (*p_ != '"') err("expected object key");
          std::string k = string();
          ws();
          if (*p_ != ':') err("expected ':'");
          ++p_;
          j->obj.emplace_back(std::move(k), value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == '}') { ++p_; break; }
          err("expected ',' or '}'");
        }
        return j;
      }
      case '[': {
        j->kind = Json::Array;
        ++p_;
        ws();
        if (*p_ == ']') { ++p_; return j; }
        while (true) {
          j->arr.push_back(value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == ']') { ++p_; break; }
          err("expected ',' or ']'");
        }
        return j;
      }
      case '"':
        j->kind = Json::String;
        j->str = string();
        return j;
      case 't':
        if (strncmp(p_, "true", 4)) err("bad literal");
        p_ += 4;
        j->kind = Json::Bool;
        j->b = true;
        return j;
      case 'f':
        if (strncmp(p_, "false", 5)) err("bad literal");
        p_ += 5;
        j->kind = Json::Bool;
        return j;
      case 'n':
        if (strncmp(p_, "null", 4)) err("bad literal");
        p_ += 4;
        return j;
      default:
        return number(j);
    }
  }
  JsonPtr number(JsonPtr j) {
    const char* s = p_;
    if (*p_ == '-') ++p_;
    if (!(*p_ >= '0' && *p_ <= '9')) err("unexpected character");
    bool integral = true;
    while ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-') {
      if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') integral = false;
      ++p_;
    }
    j->kind = Json::Number;
    std::string t(s, p_ - s);
    j->num = strtod(t.c_str(), nullptr);
    if (integral) {
      j->is_int = true;
      j->i64 = t[0] == '-' ? int64_t(strtoll(t.c_str(), nullptr, 10)) : int64_t(strtoull(t.c_str(), nullptr, 10));
    }
    return j;
  }
  std::string string() {
    std::string out;
    ++p_;  // opening quote
    while (*p_ && *p_ != '"') {
      if (*p_ == '\\') {
        ++p_;
        switch (*p_) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = 0;
            for (int i = 1; i <= 4; ++i) {
              char c = p_[i];
              cp = cp * 16 + (c >= '0' && c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
            }
            p_ += 4;
            if (cp < 0x80) out += char(cp);
            else if (cp < 0x800) { out += char(0xc0 | (cp >> 6)); out += char(0x80 | (cp & 0x3f)); }
            else { out += char(0xe0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3f)); out += char(0x80 | (cp & 0x3f)); }
            break;
          }
          default: out += *p_;
        }
        ++p_;
      } else {
        out += *p_++;
      }
    }
    if (*p_ != '"') err("unterminated string");
    ++p_;
    return out;
  }
};

}  // namespace flock
