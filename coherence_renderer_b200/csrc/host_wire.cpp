// host_wire.cpp — the front end's socket format (SURVEY.md §8f N4 meeting N1): Camlpy.marshall / unmarshall
// (camlpy.ml:18-124; the Python side of the same format is pycaml.py:30-98) and the "RefreshWindow" message
// Wxgui.refresh_window sends (wxgui.ml:352-366).  A marshallable crosses the C ABI as its pre-order token list.
// Plain C++, no device work: the pixels of a RefreshWindow message are written behind the header by
// coh_wire_refresh_window (coherence_b200.cu) straight from the GPU framebuffer.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../include/coherence_b200.h"

namespace {
// camlpy.ml:33-37: the low 32 bits, most significant byte first
inline void put_u32(uint8_t* p, uint64_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
// camlpy.ml:85-86: no sign extension (OCaml ints are wider than the word): 0 .. 2^32 - 1
inline int64_t get_u32(const uint8_t* p) { return ((int64_t)p[0] << 24) | ((int64_t)p[1] << 16) | ((int64_t)p[2] << 8) | (int64_t)p[3]; }
struct Open { int64_t len_pos; int64_t left; };   // a Tuple whose members are still being written
}  // namespace

extern "C" {

// Camlpy.marshall (camlpy.ml:39-82): 4 bytes of size, then the flattened value
int64_t coh_host_wire_marshal(const int32_t* kinds, const int64_t* values, const int64_t* offsets, int32_t n_tokens,
                              const uint8_t* strings, uint8_t* out, int64_t cap) {
  if (n_tokens <= 0) return -1;
  for (int pass = 0; pass < 2; pass++) {          // pass 0 measures (camlpy.ml:64-75), pass 1 writes
    const bool wr = pass == 1;
    int64_t pos = 4;
    std::vector<Open> open;
    for (int32_t i = 0; i < n_tokens; i++) {
      if (i > 0 && open.empty()) return -1;       // a second top-level value
      const int64_t v = values[i];
      bool member_done = true;
      switch (kinds[i]) {
        case COH_WIRE_UNIT: if (wr) out[pos] = COH_WIRE_UNIT; pos += 1; break;
        case COH_WIRE_INT: if (wr) { out[pos] = COH_WIRE_INT; put_u32(out + pos + 1, (uint64_t)v); } pos += 5; break;
        case COH_WIRE_BOOL: if (wr) { out[pos] = COH_WIRE_BOOL; out[pos + 1] = v ? 1 : 0; } pos += 2; break;
        case COH_WIRE_STRING:
          if (v < 0 || v > 0xFFFFFFFFll || (v > 0 && !strings)) return -1;
          if (wr) { out[pos] = COH_WIRE_STRING; put_u32(out + pos + 1, (uint64_t)v); if (v) memcpy(out + pos + 5, strings + offsets[i], (size_t)v); }
          pos += 5 + v; break;
        case COH_WIRE_TUPLE:
          if (v < 0) return -1;
          if (wr) { out[pos] = COH_WIRE_TUPLE; put_u32(out + pos + 1, 0); }
          if (v > 0) { open.push_back({pos + 1, v}); member_done = false; }
          pos += 5; break;
        default: return -1;
      }
      if (member_done)
        while (!open.empty() && --open.back().left == 0) {   // the last member closes its Tuple, which is a member itself
          if (wr) put_u32(out + open.back().len_pos, (uint64_t)(pos - (open.back().len_pos + 4)));
          open.pop_back();
        }
    }
    if (!open.empty()) return -1;                 // members missing
    if (!wr) { if (!out || pos > cap) return pos; }
    else { put_u32(out, (uint64_t)(pos - 4)); return pos; }
  }
  return -1;
}

// Camlpy.unmarshall (camlpy.ml:88-124): None (*taken = 0) until the whole message has arrived, Invalid_data (-1) for
// anything but exactly one well-formed value.  Offsets of Strings are positions in buf.
int32_t coh_host_wire_unmarshal(const uint8_t* buf, int64_t n, int32_t* kinds, int64_t* values, int64_t* offsets, int32_t cap_tokens,
                                int32_t* n_tokens, int64_t* taken) {
  if (n_tokens) *n_tokens = 0;
  if (taken) *taken = 0;
  if (n < 4) return 0;
  const int64_t len = get_u32(buf);
  if (n < 4 + len) return 0;
  const int64_t end = 4 + len;
  struct Up { int64_t end; int32_t tok; };
  std::vector<Up> up;                              // the Tuples around the current position
  int32_t count = 0, top_level = 0;
  auto emit = [&](int32_t kind, int64_t value, int64_t off) {
    if (count < cap_tokens) { if (kinds) kinds[count] = kind; if (values) values[count] = value; if (offsets) offsets[count] = off; }
    if (up.empty()) top_level++;
    else if (up.back().tok < cap_tokens && values) values[up.back().tok]++;
    count++;
  };
  int64_t pos = 4;
  for (;;) {
    while (!up.empty() && pos == up.back().end) up.pop_back();
    const int64_t lim = up.empty() ? end : up.back().end;
    if (pos == lim) break;                         // (only at the top level: the message is used up)
    const uint8_t t = buf[pos];
    const int64_t room = lim - pos - 1;            // bytes after the tag inside the enclosing value
    if (t == COH_WIRE_INT) { if (room < 4) return -1; emit(COH_WIRE_INT, get_u32(buf + pos + 1), 0); pos += 5; }
    else if (t == COH_WIRE_UNIT) { emit(COH_WIRE_UNIT, 0, 0); pos += 1; }
    else if (t == COH_WIRE_BOOL) { if (room < 1) return -1; emit(COH_WIRE_BOOL, buf[pos + 1] != 0, 0); pos += 2; }
    else if (t == COH_WIRE_STRING) {
      if (room < 4) return -1;
      const int64_t l = get_u32(buf + pos + 1);
      if (room - 4 < l) return -1;                 // Pdfutil.take fails on a short list
      emit(COH_WIRE_STRING, l, pos + 5); pos += 5 + l;
    } else if (t == COH_WIRE_TUPLE) {
      if (room < 4) return -1;
      const int64_t l = get_u32(buf + pos + 1);
      if (room - 4 < l) return -1;
      emit(COH_WIRE_TUPLE, 0, 0);
      up.push_back({pos + 5 + l, count - 1}); pos += 5;
    } else return -1;
  }
  if (top_level != 1) return -1;                   // camlpy.ml:119-121: [x] only
  if (n_tokens) *n_tokens = count;
  if (taken) *taken = end;
  return 0;
}

// Wxgui.refresh_window (wxgui.ml:352-366): Tuple [String "RefreshWindow"; Int window; Int xmin; Int ymin; Int width;
// Int height; String rgb888] — everything in front of the pixel bytes.  Returns the size of the whole message, 0 for the
// rectangles the reference sends nothing for (xmin = xmax or ymin = ymax), -1 where its assertion fails (wxgui.ml:335).
int64_t coh_host_wire_refresh_window(int32_t window, int32_t xmin, int32_t ymin, int32_t xmax, int32_t ymax, uint8_t header_out[64], int32_t* header_len) {
  if (header_len) *header_len = 0;
  if (xmin == xmax || ymin == ymax) return 0;
  if (xmax < xmin || ymax < ymin || xmin < 0 || ymin < 0) return -1;
  const int64_t w = (int64_t)xmax - xmin + 1, h = (int64_t)ymax - ymin + 1, data = w * h * 3;
  if (data > 0xFFFFFFFFll - 64) return -1;
  static const char name[] = "RefreshWindow";
  const int64_t name_len = (int64_t)sizeof name - 1;
  uint8_t* p = header_out;
  const int64_t body = 5 + name_len + 5 * 5 + 5 + data;          // the members of the Tuple
  put_u32(p, (uint64_t)(5 + body)); p += 4;
  *p++ = COH_WIRE_TUPLE; put_u32(p, (uint64_t)body); p += 4;
  *p++ = COH_WIRE_STRING; put_u32(p, (uint64_t)name_len); p += 4; memcpy(p, name, (size_t)name_len); p += name_len;
  const int64_t ints[5] = {window, xmin, ymin, w, h};
  for (int i = 0; i < 5; i++) { *p++ = COH_WIRE_INT; put_u32(p, (uint64_t)ints[i]); p += 4; }
  *p++ = COH_WIRE_STRING; put_u32(p, (uint64_t)data); p += 4;
  if (header_len) *header_len = (int32_t)(p - header_out);
  return (p - header_out) + data;
}

}  // extern "C"
