// ORACLE — TEST INFRASTRUCTURE ONLY.  Restatement of camlpy.ml (the front end's socket format):
// `marshallable` as a tree, `marshall` through size_of_marshallable + marshall_flatten (camlpy.ml:39-82), `unmarshall`
// through unmarshall_inner on a LIST of byte values with take / drop (camlpy.ml:84-124), as the reference writes them.
//
// PARITY PINNED (this file only): the reference holds a second implementation of the format, pycaml.py, which runs in the
// build container; tools/make_wire_golden.py imports it unedited and writes tests/golden/wire_pycaml.json, and
// tests/test_wire.py checks this restatement against those vectors (marshal and unmarshal) before using it as the checker
// of the product's coh_host_wire_* entry points.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {
namespace camlpy {

struct Marshallable {                     // camlpy.ml:3-8
  enum Kind { Tuple = 0, Unit = 1, Int = 2, String = 3, Bool = 4 } kind = Unit;   // the constructors, numbered by their tags (26-30)
  long long i = 0;                        // Int (an OCaml int: 63 bits), Bool (0 / 1)
  std::string s;                          // String
  std::vector<Marshallable> members;      // Tuple
};
struct Invalid_data : std::runtime_error { Invalid_data() : std::runtime_error("Invalid_data") {} };   // camlpy.ml:84

// camlpy.ml:33-37 bytes_of_int: (i land (255 lsl k)) lsr k for k = 24, 16, 8, 0
inline void bytes_of_int(std::string& s, size_t p, long long i) {
  s[p] = (char)((i & (255LL << 24)) >> 24); s[p + 1] = (char)((i & (255LL << 16)) >> 16);
  s[p + 2] = (char)((i & (255LL << 8)) >> 8); s[p + 3] = (char)(i & 255);
}
// camlpy.ml:39-62 marshall_flatten s pos m: returns the position behind what it wrote
inline size_t marshall_flatten(std::string& s, size_t pos, const Marshallable& m) {
  switch (m.kind) {
    case Marshallable::Unit: s[pos] = 1; return pos + 1;
    case Marshallable::Int: s[pos] = 2; bytes_of_int(s, pos + 1, m.i); return pos + 5;
    case Marshallable::Bool: s[pos] = 4; s[pos + 1] = m.i ? 1 : 0; return pos + 2;
    case Marshallable::String: {
      const size_t l = m.s.size();
      s[pos] = 3; bytes_of_int(s, pos + 1, (long long)l);
      s.replace(pos + 5, l, m.s);
      return pos + 5 + l;
    }
    default: {
      s[pos] = 0;
      size_t p = pos + 5;
      for (const Marshallable& x : m.members) p = marshall_flatten(s, p, x);
      bytes_of_int(s, pos + 1, (long long)(p - (pos + 5)));
      return p;
    }
  }
}
// camlpy.ml:64-75 size_of_marshallable (the inner function walks a work list; same sum)
inline size_t size_inner(const Marshallable& m) {
  switch (m.kind) {
    case Marshallable::Int: return 5;
    case Marshallable::Unit: return 1;
    case Marshallable::Bool: return 2;
    case Marshallable::String: return 5 + m.s.size();
    default: { size_t n = 5; for (const Marshallable& x : m.members) n += size_inner(x); return n; }
  }
}
// camlpy.ml:77-82
inline std::string marshall(const Marshallable& m) {
  const size_t size = size_inner(m) + 4;
  std::string str(size, '\0');
  bytes_of_int(str, 0, (long long)(size - 4));
  marshall_flatten(str, 4, m);
  return str;
}

typedef std::vector<int> Bytes;           // `map int_of_char (explode str)`
inline long long int_of_bytes(int i0, int i1, int i2, int i3) {   // camlpy.ml:85-86: no sign (ints are wider than 32 bits)
  return ((long long)i0 << 24) | ((long long)i1 << 16) | ((long long)i2 << 8) | (long long)i3;
}
// Pdfutil.take / drop fail on lists that are too short; the failure becomes Invalid_data (camlpy.ml:121-123)
inline Bytes take(const Bytes& l, size_t from, long long n) {
  if (n < 0 || from + (size_t)n > l.size()) throw Invalid_data();
  return Bytes(l.begin() + (long)from, l.begin() + (long)from + (long)n);
}
// camlpy.ml:88-104 unmarshall_inner: the whole list becomes a list of values
inline std::vector<Marshallable> unmarshall_inner(const Bytes& l) {
  std::vector<Marshallable> out;
  size_t p = 0;
  while (p < l.size()) {
    const int t = l[p];
    const size_t left = l.size() - p - 1;
    Marshallable m;
    if (t == 2 && left >= 4) { m.kind = Marshallable::Int; m.i = int_of_bytes(l[p + 1], l[p + 2], l[p + 3], l[p + 4]); p += 5; }
    else if (t == 1) { m.kind = Marshallable::Unit; p += 1; }
    else if (t == 4 && left >= 1) { m.kind = Marshallable::Bool; m.i = l[p + 1] != 0; p += 2; }
    else if (t == 3 && left >= 4) {
      const long long len = int_of_bytes(l[p + 1], l[p + 2], l[p + 3], l[p + 4]);
      const Bytes b = take(l, p + 5, len);
      m.kind = Marshallable::String; m.s.assign(b.begin(), b.end());
      p += 5 + (size_t)len;
    } else if (t == 0 && left >= 4) {
      const long long len = int_of_bytes(l[p + 1], l[p + 2], l[p + 3], l[p + 4]);
      m.kind = Marshallable::Tuple; m.members = unmarshall_inner(take(l, p + 5, len));
      p += 5 + (size_t)len;
    } else throw Invalid_data();
    out.push_back(m);
  }
  return out;
}
// camlpy.ml:106-124: false = None (the message has not arrived in full), true = Some (taken, value)
inline bool unmarshall(const std::string& str, long long& taken, Marshallable& value) {
  if (str.size() < 4) return false;
  const long long len = int_of_bytes((unsigned char)str[0], (unsigned char)str[1], (unsigned char)str[2], (unsigned char)str[3]);
  if (str.size() < 4 + (size_t)len) return false;
  Bytes l; l.reserve((size_t)len);
  for (long long k = 0; k < len; k++) l.push_back((unsigned char)str[4 + (size_t)k]);
  std::vector<Marshallable> vs = unmarshall_inner(l);
  if (vs.size() != 1) throw Invalid_data();
  taken = len + 4; value = vs[0];
  return true;
}

}  // namespace camlpy
}  // namespace oracle
