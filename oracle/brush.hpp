// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's brush strokes.  Follows
// /root/reference/brush.ml:14-27 (types, sizeof_brush), 60-92 (g, drawround), 102-122
// (stamp), 126-130 (points_of_brushstroke), 135-173 (shape), 176-222 (sprite), 235-331 (smear) and
// /root/reference/polygon.ml:143-218 (points_on_path).
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cmath>
#include "polygon.hpp"
#include "sprite.hpp"
namespace oracle {

// A path as the reference's Pdfgraphics.path restricted to what the rasteriser reads:
// subpaths of segments; a segment is Straight(p1,p2) or Bezier(p1..p4).
struct Segment { bool bezier; Pt p[4]; };
typedef std::vector<Segment> Subpath;
typedef std::vector<Subpath> Path;

// polygon.ml:143-218.  Beziers are flattened at curve_accuracy; flattened pieces of
// one segment are prepended as a block (segment order reversed, pieces forward).
inline double straightlength(Pt a, Pt b) {
  auto sq = [](double x) { return x * x; };
  return std::sqrt(sq(b.first - a.first) + sq(b.second - a.second));
}
inline std::vector<Pt> points_on_path(double sep, const Path& path) {
  std::vector<Pt> points;
  for (const Subpath& sub : path) {
    std::vector<std::pair<Pt, Pt>> segs;  // head of the reference list = segs.front()
    for (const Segment& s : sub) {
      if (!s.bezier) segs.insert(segs.begin(), {s.p[0], s.p[1]});
      else {
        std::vector<std::pair<Pt, Pt>> e;
        bezier_subdivide(curve_accuracy, s.p[0], s.p[1], s.p[2], s.p[3], e);
        segs.insert(segs.begin(), e.begin(), e.end());
      }
    }
    // takelength (polygon.ml:173-184): walk `sep` along the list, splitting a segment
    size_t i = 0;
    while (i < segs.size()) {
      double want = sep;
      bool found = false;
      while (i < segs.size()) {
        double l = straightlength(segs[i].first, segs[i].second);
        if (want <= l) {
          // splitat (polygon.ml:151-160)
          ORACLE_ASSERT(l > 0., "splitat: zero length");
          double prop = want / l;
          Pt p1 = segs[i].first, p2 = segs[i].second;
          Pt p(p1.first * (1. - prop) + p2.first * prop, p1.second * (1. - prop) + p2.second * prop);
          points.push_back(p);
          if (p == p2) i++; else segs[i].first = p;
          found = true;
          break;
        }
        want -= l; i++;
      }
      if (!found) break;
    }
  }
  return points;  // the reference conses then reverses: emission order
}

struct BrushStroke {
  double opacity = 1.0, radius = 1.0;       // (opacity, Gaussian radius)
  bool dummy = false; int rx = 0;           // Dummy (rx, ry), rx = ry (brush.ml:14-16; Brush.mkdummy always makes them equal)
  std::vector<std::pair<int, int>> points;  // toint(x+0.5), toint(y+0.5), in list order (brush.ml:172,196-199)
  int bw() const { return dummy ? rx * 2 + 1 : (int)std::ceil(radius) * 2 + 1; }  // brush.ml:25-28
};
inline std::vector<std::pair<int, int>> round_points(const std::vector<Pt>& pts) {
  std::vector<std::pair<int, int>> o;
  for (auto& p : pts) o.push_back({(int)(p.first + 0.5), (int)(p.second + 0.5)});
  return o;
}
// brush.ml:60-92: the (2*intr+1)^2 Gaussian stamp of `colour`.
inline std::vector<colour> drawround(double radius, double opacity, colour col, int& size) {
  ORACLE_ASSERT(radius >= 0., "drawround: radius");
  ORACLE_ASSERT(opacity >= 0. && opacity <= 1., "drawround: opacity");
  int intopacity = (int)(opacity * 255.), intr = (int)std::ceil(radius);
  size = intr * 2 + 1;
  std::vector<colour> brush((size_t)size * size);
  auto sq = [](double x) { return x * x; };
  for (int x = 1; x <= size; x++)
    for (int y = 1; y <= size; y++) {
      int xp = x - intr - 1, yp = y - intr - 1;
      double r = radius / 2.;
      double v = 255. * std::exp(-(sq((double)xp / r) + sq((double)yp / r)));
      int vi = (int)(v * 1.);
      ORACLE_ASSERT(vi >= 0 && vi <= 255, "drawround: v'");
      brush[(size_t)(y - 1) * size + (x - 1)] = dissolve(dissolve(col, intopacity), vi);
    }
  return brush;
}
// brush.ml:135-173
inline Shape shape_of_brushstroke(const BrushStroke& b) {
  std::vector<std::pair<int, int>> pts = b.points;
  std::stable_sort(pts.begin(), pts.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& c) {
    return a.second != c.second ? a.second < c.second : a.first < c.first;
  });
  Shape s;
  for (auto& p : pts) {
    if (s.rows.empty() || s.rows.back().y != p.second) s.rows.push_back({p.second, {}});
    Spanline& l = s.rows.back().spans;
    if (!l.empty() && l.back().x + l.back().len - 1 == p.first) continue;          // duplicate point
    if (!l.empty() && l.back().x + l.back().len == p.first) l.back().len++;        // abutting
    else l.push_back({p.first, 1});
  }
  int r = (b.bw() - 1) / 2;
  return bloat(r, r, s);
}
// brush.ml:176-222
inline Sprite sprite_of_brushstroke(const BrushStroke& b, const Fill& fill, const Shape& shp) {
  // brush.ml:178-181: a dummy brush is the WHOLE shape of the stroke in white, whatever fill and region are asked for
  if (b.dummy) return fillshape(shape_of_brushstroke(b), Fill::plain(mkcol(255, 255, 255)));
  Sprite none;
  if (shp.null()) return none;
  int r = (b.bw() - 1) / 2;
  Shape bloated = bloat(r, r, shp);
  Shape twice = bloat(r, r, bloated);
  Box bb; shape_bounds(twice, bb);
  Canvas canvas(bb.x0, bb.y0, bb.x1 - bb.x0 + 1, bb.y1 - bb.y0 + 1, clear_colour());
  int size;
  std::vector<colour> brush = drawround(b.radius, b.opacity, mkcol(255, 255, 255), size);
  for (auto& p : b.points) {
    if (!point_in_shape(bloated, p.first, p.second)) continue;
    for (int by = 0; by < size; by++)
      for (int bx = 0; bx < size; bx++) {
        colour& c = canvas.at(p.first - r + bx, p.second - r + by);
        c = alpha_over(c, brush[(size_t)by * size + bx]);
      }
  }
  return map_shape(shp, [&](int x, int y, int l, colour* out) {
    for (int k = 0; k < l; k++) out[k] = dissolve(fill.fillsingle(x + k, y), alpha_of_colour(canvas.at(x + k, y)));
  });
}

// ---- smearing (brush.ml:235-331; "This needs more work" in the reference: restated as it stands) ----
// brush.ml:239-257: the path subdivided until a piece's end points are at most 2 apart (Pdfutil.distance_between);
// a straight segment is subdivided as the curve p1 p p p2 with p = Pdfutil.between p1 p2 (the midpoint); the points
// are the START points of the pieces, all segments of all subpaths in order.
inline void subdivide_adjacent(Pt p1, Pt p2, Pt p3, Pt p4, std::vector<Pt>& starts) {
  auto sq = [](double x) { return x * x; };
  if (std::sqrt(sq(p1.first - p4.first) + sq(p1.second - p4.second)) <= 2.) { starts.push_back(p1); return; }
  auto half = [](Pt a, Pt b) { return Pt((a.first + b.first) / 2., (a.second + b.second) / 2.); };
  Pt l2 = half(p1, p2), h = half(p2, p3), l3 = half(l2, h), r3 = half(p3, p4), r2 = half(h, r3), l4 = half(l3, r2);
  subdivide_adjacent(p1, l2, l3, l4, starts);
  subdivide_adjacent(l4, r2, r3, p4, starts);
}
inline std::vector<Pt> points_of_brushstroke_smear(const Path& path) {
  std::vector<Pt> pts;
  for (const Subpath& sub : path)
    for (const Segment& sg : sub) {
      if (!sg.bezier) {
        Pt p((sg.p[0].first + sg.p[1].first) / 2., (sg.p[0].second + sg.p[1].second) / 2.);
        subdivide_adjacent(sg.p[0], p, p, sg.p[1], pts);
      } else subdivide_adjacent(sg.p[0], sg.p[1], sg.p[2], sg.p[3], pts);
    }
  return pts;
}
// brush.ml:259-264 drop_duplicates after toint (truncation): consecutive duplicates go
inline std::vector<std::pair<int, int>> smear_int_points(const std::vector<Pt>& pts) {
  std::vector<std::pair<int, int>> o;
  for (auto& p : pts) {
    std::pair<int, int> q((int)p.first, (int)p.second);
    if (o.empty() || o.back() != q) o.push_back(q);
  }
  return o;
}
// brush.ml:286-331 smear spr brushstroke, with the deduplicated integer points of find_smear_directions (266-283)
// given (the float path work stays in front of the boundary).  Exceptions of subcopy / stamp are swallowed per point
// like the reference's `try ... with _ -> ()`.
inline Sprite smear(const Sprite& spr0, const BrushStroke& b, const std::vector<std::pair<int, int>>& ipts) {
  // flesh the sprite out to the shape of the brush stroke
  Shape bs = shape_of_brushstroke(b);
  Sprite spr = caf(over, opaque, spr0, fillshape(bs, Fill::plain(clear_colour()))).first;
  if (spr.null()) return spr;
  const int bw = b.bw(), rad = (bw - 1) / 2;
  if (ipts.empty()) return spr;
  auto sgn = [](int x) { return x > 0 ? -1 : (x < 0 ? 1 : 0); };   // (sic)
  Box bb; shape_bounds(shape_of_sprite(spr), bb);
  const int xoff = bb.x0, yoff = bb.y0;
  // Sprite.flatten_sprite 1: a border of one pixel; canvas coordinates are 1-based
  const int CW = bb.x1 - bb.x0 + 1 + 2, CH = bb.y1 - bb.y0 + 1 + 2;
  std::vector<colour> canvas((size_t)CW * CH, clear_colour());
  auto at = [&](int cx, int cy) -> colour& { return canvas[(size_t)(cy - 1) * CW + (cx - 1)]; };
  for (auto& r : spr.rows) {
    int off = 0;
    for (auto& sp : r.spans) { for (int k = 0; k < sp.len; k++) at(sp.x + k - xoff + 2, r.y - yoff + 2) = r.px[off + k]; off += sp.len; }
  }
  int size;
  std::vector<colour> opacbrush = drawround(b.radius, b.opacity, dissolve(mkcol(255, 255, 255), 255), size);
  ORACLE_ASSERT(size == bw, "smear: brush size");
  std::vector<colour> brush((size_t)bw * bw, clear_colour());
  for (int pass = 1; pass <= 2; pass++)
    for (size_t i = 0; i < ipts.size(); i++) {
      const int dx = i ? sgn(ipts[i].first - ipts[i - 1].first) : 0, dy = i ? sgn(ipts[i].second - ipts[i - 1].second) : 0;
      const int x = ipts[i].first - xoff + 1, y = ipts[i].second - yoff + 1;
      // 1. read the brush: Canvas.subcopy canvas brush sx sy bw bw (canvas.ml:41-57)
      const int sx = x - rad + 1 - dx, sy = y - rad + 1 - dy;
      if (!(sx > 0 && sy > 0 && sx + bw - 1 <= CW && sy + bw - 1 <= CH)) continue;   // Failure "subcopy", swallowed
      for (int yd = sy; yd <= sy + bw - 1; yd++)
        for (int xd = sx; xd <= sx + bw - 1; xd++) brush[(size_t)(yd - sy) * bw + (xd - sx)] = at(xd, yd);
      // 3. stamp it at (x + 1, y + 1) (brush.ml:102-122)
      const int startx = x + 1 - rad, starty = y + 1 - rad, endx = x + 1 + rad, endy = y + 1 + rad;
      if (!(startx >= 1 && endx <= CW && starty >= 1 && endy <= CH)) continue;        // Failure "Brush.stamp", swallowed
      for (int py = starty; py <= endy; py++)
        for (int px = startx; px <= endx; px++) {
          const int bx = px - startx, by = py - starty;
          colour& a = at(px, py);
          a = dissolve_between(brush[(size_t)by * bw + bx], a, alpha_of_colour(opacbrush[(size_t)by * bw + bx]));
        }
    }
  // Sprite.pickup shp (-xoff + 3) (-yoff + 3) canvas
  return map_shape(shape_of_sprite(spr), [&](int x, int y, int l, colour* out) { for (int k = 0; k < l; k++) out[k] = at(x + k - xoff + 2, y - yoff + 2); });
}
}  // namespace oracle
