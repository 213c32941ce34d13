// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's fills and sprites (span sets with colour
// content) and of compose-and-filter.  Follows /root/reference/fill.ml:19-140 and
// /root/reference/sprite.ml:108-190 (map_shape, shapespan_iter, fillshape,
// shape_of_sprite), 307-374 (map_coords, sprite_map), 486-501 (translate), 642-721
// (portion), 730-1170 (caf), 1668-1739 (flatten_sprite / pickup).
// A sprite row keeps its canonical span list plus one colour per pixel; the
// reference's Run/Samples/Interval sub-span encodings are storage detail and are not
// modelled (pixel values and span sets are what parity is defined on).
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cmath>
#include <functional>
#include "colour.hpp"
#include "shape.hpp"
namespace oracle {

// ---- Fill (fill.ml:41-140): closures restated as a tagged record ----
struct Fill {
  enum Kind { Plain = 0, Axial = 1, Radial = 2 } kind = Plain;
  colour c0 = 0, c1 = 0;           // plain colour / cs, ce
  double p[6] = {0, 0, 0, 0, 0, 0};  // axial: x0,y0,x1,y1 ; radial: cx,cy,px,py,p'x,p'y
  bool ext_s = false, ext_e = false;
  bool fancy() const { return kind != Plain; }
  static Fill plain(colour c) { Fill f; f.kind = Plain; f.c0 = c; return f; }
  // fill.ml:77-93 / 112-127
  colour lookup(double x, double y) const {
    auto sqr = [](double v) { return v * v; };
    if (kind == Plain) return c0;
    if (kind == Axial) {
      double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
      if (x1 == x0 && y1 == y0) return clear_colour();
      double bottom = sqr(x1 - x0) + sqr(y1 - y0);
      double xp = ((x1 - x0) * (x - x0) + (y1 - y0) * (y - y0)) / bottom;
      if (xp < 0.) return ext_s ? c0 : clear_colour();
      if (xp > 1.) return ext_e ? c1 : clear_colour();
      return dissolve_between(c0, c1, 255 - (int)(xp * 255.));
    }
    auto dist = [&](double ax, double ay, double bx, double by) { return std::sqrt(sqr(ax - bx) + sqr(ay - by)); };
    double r = dist(p[0], p[1], p[2], p[3]), r2 = dist(p[0], p[1], p[4], p[5]);
    double diff = r2 - r;
    double d = dist(p[0], p[1], x, y);
    if (d > r2) return ext_e ? c1 : clear_colour();
    if (d < r) return ext_s ? c0 : clear_colour();
    if (diff == 0.) return c0;
    double t = (d - r) / diff;
    return dissolve_between(c0, c1, 255 - (int)(t * 255.));
  }
  colour fillsingle(int x, int y) const { return lookup((double)x, (double)y); }
  // fill.ml:96-102 / 130-134 — colour of pixel x+i of a span starting at x.
  void fillspan(int x, int y, int l, colour* out) const {
    for (int i = 0; i < l; i++) out[i] = lookup((double)(i + x), (double)y);
  }
};

struct SpriteRow { int y; Spanline spans; std::vector<colour> px; };
struct Sprite {
  std::vector<SpriteRow> rows;  // empty == NullSprite
  bool null() const { return rows.empty(); }
};

inline Shape shape_of_sprite(const Sprite& s) {  // sprite.ml:178-190
  Shape o;
  for (auto& r : s.rows) o.rows.push_back({r.y, r.spans});
  return o;
}
// sprite.ml:108-126 map_shape: f(x, y, len, out) fills the colours of one span.
inline Sprite map_shape(const Shape& shp, const std::function<void(int, int, int, colour*)>& f) {
  Sprite o;
  for (auto& r : shp.rows) {
    SpriteRow sr; sr.y = r.y; sr.spans = r.spans;
    int n = 0; for (auto& sp : r.spans) n += sp.len;
    sr.px.resize(n);
    int off = 0;
    for (auto& sp : r.spans) { f(sp.x, r.y, sp.len, sr.px.data() + off); off += sp.len; }
    o.rows.push_back(std::move(sr));
  }
  return o;
}
// sprite.ml:158-175 — plain fills use fillsingle 0 0, fancy ones fillspan.
inline Sprite fillshape(const Shape& shp, const Fill& fill) {
  if (fill.fancy()) return map_shape(shp, [&](int x, int y, int l, colour* out) { fill.fillspan(x, y, l, out); });
  colour c = fill.fillsingle(0, 0);
  return map_shape(shp, [&](int, int, int l, colour* out) { for (int i = 0; i < l; i++) out[i] = c; });
}
inline Sprite sprite_map(const std::function<colour(colour)>& f, const Sprite& s) {  // sprite.ml:358-374
  Sprite o = s;
  for (auto& r : o.rows) for (auto& c : r.px) c = f(c);
  return o;
}
inline Sprite translate_sprite(int dx, int dy, const Sprite& s) {  // sprite.ml:486-501
  Sprite o = s;
  for (auto& r : o.rows) { r.y += dy; for (auto& sp : r.spans) sp.x += dx; }
  return o;
}
inline long sprite_card(const Sprite& s) { long n = 0; for (auto& r : s.rows) n += (long)r.px.size(); return n; }

// sprite.ml:642-721 — restriction of a sprite to a shape that must be a subset of
// the sprite's own shape; anything else is "bad input" (Failure).
inline Sprite portion(const Sprite& spr, const Shape& shp) {
  Sprite o;
  if (shp.null()) return o;
  if (spr.null()) throw std::runtime_error("portion: malformed input (sprite null, shape not)");
  size_t i = 0;
  for (auto& r : shp.rows) {
    while (i < spr.rows.size() && spr.rows[i].y < r.y) i++;
    if (i >= spr.rows.size() || spr.rows[i].y != r.y) throw std::runtime_error("portion_vspans: bad input");
    const SpriteRow& a = spr.rows[i];
    SpriteRow sr; sr.y = r.y; sr.spans = r.spans;
    size_t k = 0; int off = 0;  // off = pixel offset of a.spans[k] in a.px
    for (auto& sp : r.spans) {
      while (k < a.spans.size() && a.spans[k].x + a.spans[k].len - 1 < sp.x) { off += a.spans[k].len; k++; }
      if (k >= a.spans.size() || a.spans[k].x > sp.x || a.spans[k].x + a.spans[k].len < sp.x + sp.len)
        throw std::runtime_error("portion_spanline: bad input");
      const colour* src = a.px.data() + off + (sp.x - a.spans[k].x);
      sr.px.insert(sr.px.end(), src, src + sp.len);
    }
    o.rows.push_back(std::move(sr));
  }
  return o;
}

// sprite.ml:889-1170 — compose-and-filter.  Result sprite covers shape(a) ∪ shape(b);
// a pixel is compop(a_px, b_px) where both are present, else the one present.  The
// returned shape is {p in shape(b) : filterop(result_px)} (b-only or overlap pixels).
// Abutting result spans fuse (spritespan_accumulate 880-884, span_accumulate 826-837).
typedef std::function<colour(colour, colour)> CompOp;
typedef std::function<bool(colour)> FilterOp;
inline void caf_row(const CompOp& compop, const FilterOp& filterop, const SpriteRow* a, const SpriteRow* b,
                    int y, Sprite& out, Shape& fout) {
  static const Spanline none;
  const Spanline& sa = a ? a->spans : none;
  const Spanline& sb = b ? b->spans : none;
  SpriteRow sr; sr.y = y;
  Spanline fs;
  size_t i = 0, j = 0; int offa = 0, offb = 0;  // pixel offsets of sa[i], sb[j]
  auto fpush = [&](int x) {
    if (!fs.empty() && fs.back().x + fs.back().len == x) fs.back().len++;
    else fs.push_back({x, 1});
  };
  auto opush = [&](int x, colour c) {
    if (!sr.spans.empty() && sr.spans.back().x + sr.spans.back().len == x) sr.spans.back().len++;
    else sr.spans.push_back({x, 1});
    sr.px.push_back(c);
  };
  while (i < sa.size() || j < sb.size()) {
    int xa = i < sa.size() ? sa[i].x : INT_MAX, xb = j < sb.size() ? sb[j].x : INT_MAX;
    int x = std::min(xa, xb);
    // advance pixel by pixel through the run starting at x until both spans are left behind
    for (;;) {
      bool ina = i < sa.size() && x >= sa[i].x && x < sa[i].x + sa[i].len;
      bool inb = j < sb.size() && x >= sb[j].x && x < sb[j].x + sb[j].len;
      if (!ina && !inb) break;
      colour c;
      if (ina && inb) c = compop(a->px[offa + (x - sa[i].x)], b->px[offb + (x - sb[j].x)]);
      else if (ina) c = a->px[offa + (x - sa[i].x)];
      else c = b->px[offb + (x - sb[j].x)];
      opush(x, c);
      if (inb && filterop(c)) fpush(x);
      x++;
      if (i < sa.size() && x == sa[i].x + sa[i].len) { offa += sa[i].len; i++; }
      if (j < sb.size() && x == sb[j].x + sb[j].len) { offb += sb[j].len; j++; }
    }
  }
  if (!sr.spans.empty()) out.rows.push_back(std::move(sr));
  if (!fs.empty()) fout.rows.push_back({y, std::move(fs)});
}
inline std::pair<Sprite, Shape> caf(const CompOp& compop, const FilterOp& filterop, const Sprite& a, const Sprite& b) {
  Sprite out; Shape fout;
  if (b.null()) return {a, fout};  // sprite.ml:1148
  size_t i = 0, j = 0;
  while (i < a.rows.size() || j < b.rows.size()) {
    int ya = i < a.rows.size() ? a.rows[i].y : INT_MAX, yb = j < b.rows.size() ? b.rows[j].y : INT_MAX;
    if (ya < yb) { out.rows.push_back(a.rows[i]); i++; }
    else if (yb < ya) { caf_row(compop, filterop, nullptr, &b.rows[j], yb, out, fout); j++; }
    else { caf_row(compop, filterop, &a.rows[i], &b.rows[j], ya, out, fout); i++; j++; }
  }
  return {std::move(out), std::move(fout)};
}

// ---- canvas round trip for brush / convolve (sprite.ml:1668-1739, canvas.ml) ----
struct Canvas {
  int x0, y0, w, h;  // covers pixels [x0, x0+w) x [y0, y0+h)
  std::vector<colour> px;
  Canvas(int x0_, int y0_, int w_, int h_, colour fill) : x0(x0_), y0(y0_), w(w_), h(h_), px((size_t)w_ * h_, fill) {}
  colour& at(int x, int y) { return px[(size_t)(y - y0) * w + (x - x0)]; }
  colour at(int x, int y) const { return px[(size_t)(y - y0) * w + (x - x0)]; }
  bool inside(int x, int y) const { return x >= x0 && x < x0 + w && y >= y0 && y < y0 + h; }
};
inline void flatten_sprite(const Sprite& s, Canvas& c) {
  for (auto& r : s.rows) {
    int off = 0;
    for (auto& sp : r.spans) { for (int k = 0; k < sp.len; k++) c.at(sp.x + k, r.y) = r.px[off + k]; off += sp.len; }
  }
}
inline Sprite pickup(const Shape& shp, const Canvas& c) {
  return map_shape(shp, [&](int x, int y, int l, colour* out) { for (int k = 0; k < l; k++) out[k] = c.at(x + k, y); });
}
}  // namespace oracle
