// Minimal JSON reader/writer for the network plan interchange format (net.cu: spn_net_to_json /
// spn_net_from_json).  Numbers are kept as text so 64-bit integers and %.17g doubles round-trip exactly.
#pragma once
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace spn {

struct JVal {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  std::string s;  // Num (text) or Str
  std::vector<JVal> a;
  std::vector<std::pair<std::string, JVal>> o;

  const JVal* find(const std::string& k) const {
    for (auto& kv : o)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const JVal& at(const std::string& k) const {
    const JVal* v = find(k);
    if (!v) throw std::runtime_error("plan json: missing key \"" + k + "\"");
    return *v;
  }
  double num() const {
    if (kind != Num) throw std::runtime_error("plan json: number expected");
    return std::strtod(s.c_str(), nullptr);
  }
  long long i64() const {
    if (kind != Num) throw std::runtime_error("plan json: integer expected");
    return std::strtoll(s.c_str(), nullptr, 10);
  }
  unsigned long long u64() const {
    if (kind != Num) throw std::runtime_error("plan json: integer expected");
    return std::strtoull(s.c_str(), nullptr, 10);
  }
  const std::string& str() const {
    if (kind != Str) throw std::runtime_error("plan json: string expected");
    return s;
  }
  const std::vector<JVal>& arr() const {
    if (kind != Arr) throw std::runtime_error("plan json: array expected");
    return a;
  }
};

class JParser {
 public:
  explicit JParser(const char* text) : p_(text) {}
  JVal parse() {
    JVal v = value();
    ws();
    if (*p_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("plan json: ") + m); }
  void ws() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
  }
  JVal value() {
    ws();
    JVal v;
    if (*p_ == '{') {
      v.kind = JVal::Obj;
      ++p_;
      ws();
      if (*p_ == '}') {
        ++p_;
        return v;
      }
      for (;;) {
        ws();
        JVal k = string();
        ws();
        if (*p_++ != ':') fail("':' expected");
        v.o.emplace_back(k.s, value());
        ws();
        if (*p_ == ',') {
          ++p_;
          continue;
        }
        if (*p_ == '}') {
          ++p_;
          return v;
        }
        fail("',' or '}' expected");
      }
    }
    if (*p_ == '[') {
      v.kind = JVal::Arr;
      ++p_;
      ws();
      if (*p_ == ']') {
        ++p_;
        return v;
      }
      for (;;) {
        v.a.push_back(value());
        ws();
        if (*p_ == ',') {
          ++p_;
          continue;
        }
        if (*p_ == ']') {
          ++p_;
          return v;
        }
        fail("',' or ']' expected");
      }
    }
    if (*p_ == '"') return string();
    if (!std::string(p_, 4).compare("true")) {
      p_ += 4;
      v.kind = JVal::Bool;
      v.b = true;
      return v;
    }
    if (!std::string(p_, 5).compare("false")) {
      p_ += 5;
      v.kind = JVal::Bool;
      return v;
    }
    if (!std::string(p_, 4).compare("null")) {
      p_ += 4;
      return v;
    }
    if (!std::string(p_, 3).compare("NaN")) {   // written by jlit() for a NaN constant
      p_ += 3;
      v.kind = JVal::Num;
      v.s = "nan";
      return v;
    }
    const char* q = p_;
    while (*q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E' || (*q >= '0' && *q <= '9')) ++q;
    if (q == p_) fail("value expected");
    v.kind = JVal::Num;
    v.s.assign(p_, q);
    p_ = q;
    return v;
  }
  JVal string() {
    if (*p_ != '"') fail("string expected");
    ++p_;
    JVal v;
    v.kind = JVal::Str;
    while (*p_ && *p_ != '"') {
      if (*p_ == '\\' && p_[1]) ++p_;
      v.s.push_back(*p_++);
    }
    if (*p_++ != '"') fail("unterminated string");
    return v;
  }
};

}  // namespace spn
