// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's 31-bit premultiplied RGBA colour codec and
// Porter–Duff operators.  Follows /root/reference/colour.ml:66-398.
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cstdint>
#include <stdexcept>
namespace oracle {

typedef int32_t colour;  // colour.ml:13 `type colour = int` (31 significant bits)

struct Nocover : std::runtime_error {  // colour.ml:21
  Nocover() : std::runtime_error("Colour.Nocover") {}
};
struct AssertFailure : std::runtime_error {
  explicit AssertFailure(const char* w) : std::runtime_error(w) {}
};
#define ORACLE_ASSERT(c, w) do { if (!(c)) throw ::oracle::AssertFailure(w); } while (0)

// colour.ml:66-79 bit masks.
constexpr int mask_equality = 1 << 30;
constexpr int mask_r_lsb = 1 << 29;
constexpr int mask_g_lsb = 1 << 28;
constexpr int mask_channel3 = 0x7F << 21;
constexpr int mask_channel2 = 0x7F << 14;
constexpr int mask_channel1 = 0x7F << 7;
constexpr int mask_channel0 = 0x7F;
constexpr int mask_b_lsb = 1 << 27;
constexpr int mask_a_lsb = 1 << 26;
constexpr int mask_r_eq_a = 1 << 25;
constexpr int mask_g_eq_a = 1 << 24;
constexpr int mask_b_eq_a = 1 << 23;

inline int concat4(int r, int g, int b, int a) {  // colour.ml:82-83
  return (r << 21) | (g << 14) | (b << 7) | a;
}
inline int index_max4(int a, int b, int c, int d) {  // colour.ml:86-96
  if (a > b) {
    if (c > d) return a > c ? 0 : 2;
    return a > d ? 0 : 3;
  }
  if (c > d) return b > c ? 1 : 2;
  return b > d ? 1 : 3;
}

// colour.ml:99-132.  `let ... and ...` binds simultaneously: the *_lsb tests see
// the ORIGINAL 8-bit channel values, the comparisons see the 7-bit halves.
inline colour colour_of_rgba(int r8, int g8, int b8, int a8) {
  int r = r8 >> 1, g = g8 >> 1, b = b8 >> 1, a = a8 >> 1;
  bool r_lsb = r8 & 1, g_lsb = g8 & 1, b_lsb = b8 & 1, a_lsb = a8 & 1;
  if (r != a && g != a && b != a) {
    return (r_lsb ? mask_r_lsb : 0) | (g_lsb ? mask_g_lsb : 0) |
           (b_lsb ? (a_lsb ? concat4(r, g, b, a) : concat4(r, g, a, b))
                  : (a_lsb ? concat4(r, a, b, g) : concat4(a, g, b, r)));
  }
  int tail;
  if (r == a) tail = concat4(0, g, b, a);
  else if (g == a) tail = concat4(0, r, b, a);
  else { ORACLE_ASSERT(b == a, "colour_of_rgba"); tail = concat4(0, r, g, a); }
  return mask_equality | (r_lsb ? mask_r_lsb : 0) | (g_lsb ? mask_g_lsb : 0) |
         (b_lsb ? mask_b_lsb : 0) | (a_lsb ? mask_a_lsb : 0) |
         (r == a ? mask_r_eq_a : 0) | (g == a ? mask_g_eq_a : 0) |
         (b == a ? mask_b_eq_a : 0) | tail;
}
inline int unsplit(int i, bool lsb) { return (i << 1) | (lsb ? 1 : 0); }  // colour.ml:134

// colour.ml:138-172
inline void rgba_of_colour(colour c, int& R, int& G, int& B, int& A) {
  int r = 0, g = 0, b = 0, a = 0;
  bool r_lsb = (mask_r_lsb & c) != 0, g_lsb = (mask_g_lsb & c) != 0;
  bool b_lsb = false, a_lsb = false;
  if ((c & mask_equality) == 0) {
    int c3 = (c & mask_channel3) >> 21, c2 = (c & mask_channel2) >> 14;
    int c1 = (c & mask_channel1) >> 7, c0 = c & mask_channel0;
    switch (index_max4(c3, c2, c1, c0)) {
      case 3: b_lsb = true; a_lsb = true; r = c3; g = c2; b = c1; a = c0; break;
      case 2: b_lsb = true; a_lsb = false; r = c3; g = c2; a = c1; b = c0; break;
      case 1: b_lsb = false; a_lsb = true; r = c3; a = c2; b = c1; g = c0; break;
      default: b_lsb = false; a_lsb = false; a = c3; g = c2; b = c1; r = c0; break;
    }
  } else {
    b_lsb = (mask_b_lsb & c) != 0;
    a_lsb = (mask_a_lsb & c) != 0;
    int c2 = (c & mask_channel2) >> 14, c1 = (c & mask_channel1) >> 7, c0 = c & mask_channel0;
    a = c0;
    if (c & mask_r_eq_a) { r = a; g = c2; b = c1; }
    else if (c & mask_g_eq_a) { g = a; r = c2; b = c1; }
    else { ORACLE_ASSERT(c & mask_b_eq_a, "rgba_of_colour"); b = a; r = c2; g = c1; }
  }
  R = unsplit(r, r_lsb); G = unsplit(g, g_lsb); B = unsplit(b, b_lsb); A = unsplit(a, a_lsb);
}
// colour.ml:174-244 — the reference's per-channel decoders are separate code but pick
// the same channel slot / lsb as rgba_of_colour in every branch (checked by reading
// both); they are restated through it.
inline int red_of_colour(colour c) { int r, g, b, a; rgba_of_colour(c, r, g, b, a); return r; }
inline int green_of_colour(colour c) { int r, g, b, a; rgba_of_colour(c, r, g, b, a); return g; }
inline int blue_of_colour(colour c) { int r, g, b, a; rgba_of_colour(c, r, g, b, a); return b; }
inline int alpha_of_colour(colour c) { int r, g, b, a; rgba_of_colour(c, r, g, b, a); return a; }

// colour.ml:247-252 — int_of_float truncation.
inline colour colour_of_rgba_float(double r, double g, double b, double a) {
  ORACLE_ASSERT(r >= 0. && g >= 0. && b >= 0. && a >= 0., "colour_of_rgba_float");
  ORACLE_ASSERT(r <= 1. && g <= 1. && b <= 1. && a <= 1., "colour_of_rgba_float");
  return colour_of_rgba((int)(r * 255.), (int)(g * 255.), (int)(b * 255.), (int)(a * 255.));
}
inline colour colour_of_channel(int a) { return colour_of_rgba(a, a, a, a); }  // colour.ml:259
inline colour clear_colour() { return colour_of_rgba(0, 0, 0, 0); }            // colour.ml:263
inline colour mkcol(int r, int g, int b) { return colour_of_rgba(r, g, b, 255); }

// colour.ml:266-280
inline colour red_channel(colour c) { return colour_of_rgba(red_of_colour(c), 0, 0, alpha_of_colour(c)); }
inline colour green_channel(colour c) { return colour_of_rgba(0, green_of_colour(c), 0, alpha_of_colour(c)); }
inline colour blue_channel(colour c) { return colour_of_rgba(0, 0, blue_of_colour(c), alpha_of_colour(c)); }
inline colour monochrome(colour c) {
  int r, g, b, a; rgba_of_colour(c, r, g, b, a);
  int av = (r + g + b) / 3;
  return colour_of_rgba(av, av, av, a);
}

inline int div255(int i) { return (i + (i >> 8) + 1) >> 8; }  // colour.ml:287

// colour.ml:291-304
inline colour dissolve(colour col, int delta) {
  ORACLE_ASSERT(delta >= 0 && delta <= 255, "dissolve: delta");
  if (delta == 0) return clear_colour();
  if (delta == 255) return col;
  int r, g, b, a; rgba_of_colour(col, r, g, b, a);
  return colour_of_rgba(div255(r * delta), div255(g * delta), div255(b * delta), div255(a * delta));
}
// colour.ml:310-311
inline int prelerp(int p, int q, int a) {
  int t = a * p + 128;
  return p + q - (((t >> 8) + t) >> 8);
}
// colour.ml:314-328
inline colour over(colour a, colour b) {
  int ra, ga, ba, aa; rgba_of_colour(a, ra, ga, ba, aa);
  if (aa == 0) return b;
  if (aa == 255) return a;
  int rb, gb, bb, ab; rgba_of_colour(b, rb, gb, bb, ab);
  return colour_of_rgba(prelerp(rb, ra, aa), prelerp(gb, ga, aa), prelerp(bb, ba, aa), prelerp(ab, aa, aa));
}
// colour.ml:332-336
inline colour alpha_over(colour a, colour b) {
  int aa = alpha_of_colour(a);
  if (aa == 0) return b;
  if (aa == 255) return a;
  int ab = alpha_of_colour(b);
  return colour_of_rgba(0, 0, 0, prelerp(ab, aa, aa));
}
// colour.ml:339-352
inline colour pd_plus(colour a, colour b) {
  int ar, ag, ab, aa, br, bg, bb, ba;
  rgba_of_colour(a, ar, ag, ab, aa); rgba_of_colour(b, br, bg, bb, ba);
  ORACLE_ASSERT(ar + br <= 255 && ag + bg <= 255 && ab + bb <= 255 && aa + ba <= 255, "pd_plus");
  return colour_of_rgba(ar + br, ag + bg, ab + bb, aa + ba);
}
// colour.ml:355-361
inline colour dissolve_between(colour a, colour b, int alpha) {
  ORACLE_ASSERT(alpha >= 0 && alpha <= 255, "dissolve_between");
  if (alpha == 0) return b;
  if (alpha == 255) return a;
  return pd_plus(dissolve(a, alpha), dissolve(b, 255 - alpha));
}
inline bool opaque(colour c) { return alpha_of_colour(c) == 255; }      // colour.ml:364
inline bool transparent(colour c) { return alpha_of_colour(c) == 0; }   // colour.ml:367
inline colour nocover(colour, colour) { throw Nocover(); }              // colour.ml:24

// Boundary helpers (not in the reference): RGBA8 word <-> colour, little-endian r,g,b,a.
inline uint32_t rgba8_of_colour(colour c) {
  int r, g, b, a; rgba_of_colour(c, r, g, b, a);
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
}
inline colour colour_of_rgba8(uint32_t w) {
  return colour_of_rgba(w & 255, (w >> 8) & 255, (w >> 16) & 255, (w >> 24) & 255);
}
}  // namespace oracle
