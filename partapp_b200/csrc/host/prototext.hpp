// prototext.hpp -- protobuf *text format* parser and proto2 wire writer for the three messages the path touches.
//
// The reference parses its experiment files with protobuf 2.3 TextFormat (libProtoBuf/protobuf_aux.hpp:56-70:
// parse_message_from_text_file) and writes HypothesisList with SerializeToOstream (:72-82).  libprotobuf is not
// available here, so this is a small generic text-format reader (scalars, strings, nested messages with `{}` or `<>`,
// repeated fields, `#` comments) plus the handful of wire-format primitives HypothesisList needs.
#pragma once

#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace prototext {

struct Node {
  // field name -> occurrences in file order; an occurrence is a scalar token or a nested message
  struct Value {
    std::string scalar;
    std::shared_ptr<Node> msg;
  };
  std::map<std::string, std::vector<Value>> fields;

  bool has(const std::string &k) const { return fields.count(k) && !fields.at(k).empty(); }
  size_t count(const std::string &k) const { return fields.count(k) ? fields.at(k).size() : 0; }
  const Value &get(const std::string &k, size_t i = 0) const { return fields.at(k).at(i); }
  std::string str(const std::string &k, const std::string &def = "") const { return has(k) ? get(k).scalar : def; }
  double num(const std::string &k, double def) const { return has(k) ? atof(get(k).scalar.c_str()) : def; }
  bool boolean(const std::string &k, bool def) const {
    if (!has(k)) return def;
    const std::string &s = get(k).scalar;
    return s == "true" || s == "1" || s == "True" || s == "t";
  }
  const Node &msg(const std::string &k, size_t i) const { return *fields.at(k).at(i).msg; }
};

class Parser {
 public:
  explicit Parser(const std::string &text) : s_(text) {}
  Node parse() {
    Node n;
    parse_fields(n, '\0');
    return n;
  }

 private:
  void skip() {
    for (;;) {
      while (i_ < s_.size() && isspace((unsigned char)s_[i_])) ++i_;
      if (i_ < s_.size() && s_[i_] == '#') {
        while (i_ < s_.size() && s_[i_] != '\n') ++i_;
      } else {
        break;
      }
    }
  }
  std::string ident() {
    size_t b = i_;
    while (i_ < s_.size() && (isalnum((unsigned char)s_[i_]) || s_[i_] == '_' || s_[i_] == '.')) ++i_;
    if (b == i_) throw std::runtime_error("prototext: expected a field name at offset " + std::to_string(b));
    return s_.substr(b, i_ - b);
  }
  std::string quoted() {
    char q = s_[i_++];
    std::string out;
    while (i_ < s_.size() && s_[i_] != q) {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        char c = s_[++i_];
        switch (c) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case '\\': out += '\\'; break;
          case '"': out += '"'; break;
          case '\'': out += '\''; break;
          default: out += c;
        }
        ++i_;
      } else {
        out += s_[i_++];
      }
    }
    if (i_ >= s_.size()) throw std::runtime_error("prototext: unterminated string");
    ++i_;
    return out;
  }
  void parse_fields(Node &n, char closer) {
    for (;;) {
      skip();
      if (i_ >= s_.size()) {
        if (closer) throw std::runtime_error("prototext: missing closing brace");
        return;
      }
      if (closer && s_[i_] == closer) {
        ++i_;
        return;
      }
      std::string key = ident();
      skip();
      Node::Value v;
      if (i_ < s_.size() && s_[i_] == ':') {
        ++i_;
        skip();
      }
      if (i_ < s_.size() && (s_[i_] == '{' || s_[i_] == '<')) {
        char close = s_[i_] == '{' ? '}' : '>';
        ++i_;
        v.msg = std::make_shared<Node>();
        parse_fields(*v.msg, close);
      } else if (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) {
        v.scalar = quoted();
        skip();
        while (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) {  // adjacent string literals concatenate
          v.scalar += quoted();
          skip();
        }
      } else {
        size_t b = i_;
        while (i_ < s_.size() && !isspace((unsigned char)s_[i_]) && s_[i_] != '}' && s_[i_] != '>' && s_[i_] != '#' &&
               s_[i_] != ';' && s_[i_] != ',')
          ++i_;
        v.scalar = s_.substr(b, i_ - b);
      }
      skip();
      if (i_ < s_.size() && (s_[i_] == ';' || s_[i_] == ',')) ++i_;
      n.fields[key].push_back(v);
    }
  }
  const std::string &s_;
  size_t i_ = 0;
};

inline Node parse_file(const std::string &path) {
  std::ifstream f(path.c_str());
  if (!f) throw std::runtime_error("prototext: cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  std::string text = ss.str();
  return Parser(text).parse();
}

// ---- proto2 wire format (HypothesisList.proto:1-10) ---------------------------------------------------------------
namespace wire {
inline void varint(std::string &b, uint64_t v) {
  while (v >= 0x80) {
    b += (char)((v & 0x7f) | 0x80);
    v >>= 7;
  }
  b += (char)v;
}
inline void key(std::string &b, int field, int type) { varint(b, (uint64_t)(field << 3 | type)); }
inline void f32(std::string &b, int field, float v) {
  key(b, field, 5);
  char t[4];
  memcpy(t, &v, 4);
  b.append(t, 4);
}
inline void boolean(std::string &b, int field, bool v) {
  key(b, field, 0);
  varint(b, v ? 1 : 0);
}
inline void bytes(std::string &b, int field, const std::string &payload) {
  key(b, field, 2);
  varint(b, payload.size());
  b += payload;
}
}  // namespace wire

}  // namespace prototext
