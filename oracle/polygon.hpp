// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's polygon scan converter and its correlated-matte
// antialiasing.  Follows /root/reference/polygon.ml:19-26 (constants), 79-127 (edges,
// bezier flattening), 235-244 (edge projections, sort), 326-388 (crossings, band
// clipping), 394-401 (spanacc), 444-528 (coverage, winding spans, shape/minshape of a
// row), 532-609 (gradient, row loop, assembly), 616-750 (AA tables, scaled shape,
// pixel_coverage, polygon_sprite).
// The list-based row loop is restated with vectors in the same order (active list =
// survivors @ newly active; crossings consed then stably sorted).
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cmath>
#include <utility>
#include "colour.hpp"
#include "coord.hpp"
#include "shape.hpp"
namespace oracle {

constexpr double curve_accuracy = 0.2;  // polygon.ml:19
constexpr int aa_res = 32;              // polygon.ml:22
constexpr double aa_softness = 2.0;     // polygon.ml:26

struct Edge { int x0, x1, y0, y1; };    // polygon.ml:79 (sub-bins)
enum Winding { NonZero = 0, EvenOdd = 1 };

typedef std::pair<double, double> Pt;

// polygon.ml:83-90
inline double distance_point_from_line(Pt c, Pt a, Pt b) {
  auto square = [](double x) { return x * x; };
  double l = std::sqrt(square(b.first - a.first) + square(b.second - a.second));
  double s = ((a.second - c.second) * (b.first - a.first) - (a.first - c.first) * (b.second - a.second)) / square(l);
  return std::fabs(s) * l;
}
// polygon.ml:107-114 — classify_float FP_normal: finite, non-zero, not subnormal.
inline bool bezier_epsilon(double eps, Pt p1, Pt p2, Pt p3, Pt p4) {
  double d1 = distance_point_from_line(p2, p1, p4), d2 = distance_point_from_line(p3, p1, p4);
  if (std::fpclassify(d1) == FP_NORMAL && std::fpclassify(d2) == FP_NORMAL) return d1 < eps && d2 < eps;
  return true;
}
// polygon.ml:119-127
inline void bezier_subdivide(double eps, Pt p1, Pt p2, Pt p3, Pt p4, std::vector<std::pair<Pt, Pt>>& out) {
  if (bezier_epsilon(eps, p1, p2, p3, p4)) { out.push_back({p1, p4}); return; }
  auto half = [](Pt a, Pt b) { return Pt((a.first + b.first) / 2., (a.second + b.second) / 2.); };
  Pt l2 = half(p1, p2), h = half(p2, p3);
  Pt l3 = half(l2, h), r3 = half(p3, p4);
  Pt r2 = half(h, r3);
  Pt l4 = half(l3, r2);
  bezier_subdivide(eps, p1, l2, l3, l4, out);
  bezier_subdivide(eps, l4, r2, r3, p4, out);
}

// polygon.ml:235-240
inline int x0in(const Edge& e) { return e.y0 > e.y1 ? e.x1 : (e.y1 > e.y0 ? e.x0 : std::min(e.x0, e.x1)); }
inline int x1in(const Edge& e) { return e.y0 > e.y1 ? e.x0 : (e.y1 > e.y0 ? e.x1 : std::max(e.x0, e.x1)); }
inline int xminin(const Edge& e) { return std::min(e.x0, e.x1); }
inline int xmaxin(const Edge& e) { return std::max(e.x0, e.x1); }
inline int yminin(const Edge& e) { return std::min(e.y0, e.y1); }
inline int ymaxin(const Edge& e) { return std::max(e.y0, e.y1); }

// polygon.ml:243-244 — stable, descending maximum y.
inline void sort_edgelist_maxy_rev(std::vector<Edge>& es) {
  std::stable_sort(es.begin(), es.end(), [](const Edge& a, const Edge& b) { return ymaxin(a) > ymaxin(b); });
}

struct Crossing { int pos; int dir; };  // polygon.ml:312-318: A = +1, C = -1
inline Crossing crossing_of_line(const Edge& e, int n) { return {n, e.y1 > e.y0 ? 1 : -1}; }  // polygon.ml:325-328

struct GEdge { double g; Edge e; };
// polygon.ml:532-535
inline GEdge gradient(const Edge& e) {
  int denom = ymaxin(e) - yminin(e);
  if (denom == 0) return {0., e};
  return {(double)(x1in(e) - x0in(e)) / (double)denom, e};
}

// polygon.ml:332-388.  Lists are consed in the reference (reverse order of the walk);
// callers reverse before the stable sort to keep the same tie order.
inline void clip_yrange_points(int top, int bot, const std::vector<GEdge>& edges,
                               std::vector<Crossing>& tops, std::vector<Edge>& middles,
                               std::vector<Crossing>& bots) {
  for (const GEdge& ge : edges) {
    const Edge& edge = ge.e;
    double g = ge.g;
    int x0 = x0in(edge), x1 = x1in(edge), ymin = yminin(edge), ymax = ymaxin(edge);
    if (ymin > bot || ymax < top) continue;
    if (ymin == ymax) { middles.push_back(edge); continue; }
    if (ymin >= top && ymax <= bot) { middles.push_back(edge); continue; }
    if (ymin >= top) {  // just bottom clipping
      int y = bot;
      int xy = (int)((double)x0 + g * ((double)(y - ymin) + 0.25) + 0.5);
      middles.push_back(Edge{x0, xy, ymin, y});
      bots.push_back(crossing_of_line(edge, xy));
    } else if (ymax <= bot) {  // just top clipping
      int y = top - 1;
      int xy = (int)((double)x0 + g * ((double)(y - ymin) + 0.25) + 0.5);
      middles.push_back(Edge{xy, x1, y + 1, ymax});
      tops.push_back(crossing_of_line(edge, xy));
    } else {  // clip both: the bottom crossing restarts from the ROUNDED top crossing
      int y = top - 1;
      int topcrossing = (int)((double)x0 + g * ((double)(y - ymin) + 0.25) + 0.5);
      Edge e2{topcrossing, x1, y + 1, ymax};
      int yb = bot;
      int x0b = x0in(e2), yminb = yminin(e2);
      int botcrossing = (int)((double)x0b + g * ((double)(yb - yminb) + 0.25) + 0.5);
      middles.push_back(Edge{x0b, botcrossing, yminb, yb});
      tops.push_back(crossing_of_line(edge, topcrossing));
      bots.push_back(crossing_of_line(edge, botcrossing));
    }
  }
  std::reverse(tops.begin(), tops.end());
  std::reverse(middles.begin(), middles.end());
  std::reverse(bots.begin(), bots.end());
}

// polygon.ml:394-401.  `spans` is kept in increasing order (the reference conses and
// reverses at the end).
inline void spanacc(Spanline& spans, int s, int e) {
  if (spans.empty()) { spans.push_back({s, e - s + 1}); return; }
  Span& last = spans.back();
  int olde = last.x + last.len - 1;
  if (s > olde + 1) spans.push_back({s, e - s + 1});
  else if (e <= olde) return;
  else last.len = e - last.x + 1;
}
// polygon.ml:444-453
inline Spanline coverage(std::vector<Edge> edgelist) {
  std::stable_sort(edgelist.begin(), edgelist.end(), [](const Edge& a, const Edge& b) { return xminin(a) < xminin(b); });
  Spanline spans;
  for (const Edge& e : edgelist) spanacc(spans, pix_of_sub(xminin(e) - halfips), pix_of_sub(xmaxin(e) + halfips));
  return spans;
}
inline void sort_crossings(std::vector<Crossing>& p) {
  std::stable_sort(p.begin(), p.end(), [](const Crossing& a, const Crossing& b) { return a.pos < b.pos; });
}
// polygon.ml:456-479 (even-odd; aa = un-widened variant)
inline Spanline spans_of_edgepoints(std::vector<Crossing> p, bool aa) {
  sort_crossings(p);
  Spanline spans;
  for (size_t i = 0; i + 1 < p.size(); i += 2) {
    int s = aa ? pix_of_sub(p[i].pos) : pix_of_sub(p[i].pos - halfips);
    int e = aa ? pix_of_sub(p[i + 1].pos) : pix_of_sub(p[i + 1].pos + halfips);
    spanacc(spans, s, e);
  }
  return spans;
}
// polygon.ml:482-512 (non-zero; aa = un-widened variant)
inline Spanline nonzero_findspans(std::vector<Crossing> p, bool aa) {
  sort_crossings(p);
  Spanline spans;
  int c = 0;
  for (size_t i = 0; i + 1 < p.size(); i++) {
    c += p[i].dir;
    if (c != 0) {
      int s = aa ? pix_of_sub(p[i].pos) : pix_of_sub(p[i].pos - halfips);
      int e = aa ? pix_of_sub(p[i + 1].pos) : pix_of_sub(p[i + 1].pos + halfips);
      spanacc(spans, s, e);
    }
  }
  return spans;
}
inline Spanline findspans(const std::vector<Crossing>& p, Winding w, bool aa) {
  return w == NonZero ? nonzero_findspans(p, aa) : spans_of_edgepoints(p, aa);
}
// polygon.ml:520-528
inline void shapeminshape_spanline(const std::vector<Crossing>& tops, const std::vector<Edge>& middles,
                                   const std::vector<Crossing>& bots, Winding w, bool aa,
                                   Spanline& line_shp, Spanline& line_minshp) {
  Spanline t = findspans(tops, w, aa), b = findspans(bots, w, aa);
  Spanline c = coverage(middles);
  Spanline tb = spanline_union(t, b);
  line_shp = spanline_union(tb, c);
  line_minshp = spanline_difference(line_shp, c);
}

// polygon.ml:538-603: polygon_spanline (row loop, y descending) + polygon +
// recompress_vspans + boxshape.  `edges` must be sorted by descending max y
// (the reference takes the first edge's ymax as the start row).
inline void shapeminshape_of_edgelist(const std::vector<Edge>& edges, Winding w, bool aa,
                                      Shape& shape, Shape& minshape) {
  shape.rows.clear(); minshape.rows.clear();
  if (edges.empty()) return;
  int y = pix_of_sub(ymaxin(edges[0]) + halfips);  // polygon.ml:564
  size_t mel = 0;                                   // index of the first not-yet-active edge
  std::vector<GEdge> ael;
  std::vector<ShapeRow> lines, lines_ms;            // collected in descending y
  for (;;) {
    int top = left_of_pix(y) - halfips;             // polygon.ml:539
    int bottom = top + 2 * ipspacing - 1;           // polygon.ml:540
    bool mel_was_empty = mel >= edges.size();
    std::vector<GEdge> ael2;
    for (const GEdge& ge : ael) if (!(yminin(ge.e) > bottom)) ael2.push_back(ge);  // lose, 546
    while (mel < edges.size() && ymaxin(edges[mel]) >= top) ael2.push_back(gradient(edges[mel++]));  // 542-546
    if (mel_was_empty && ael2.empty()) break;       // polygon.ml:548-550
    std::vector<Crossing> tops, bots; std::vector<Edge> middles;
    clip_yrange_points(top, bottom, ael2, tops, middles, bots);
    Spanline ls, lm;
    shapeminshape_spanline(tops, middles, bots, w, aa, ls, lm);
    lines.push_back({y, std::move(ls)});
    lines_ms.push_back({y, std::move(lm)});
    ael.swap(ael2);
    y--;
  }
  // recompress_vspans (571-580): rows with no spans split vspans, i.e. are dropped here.
  for (size_t i = lines.size(); i-- > 0;) {
    if (!lines[i].spans.empty()) shape.rows.push_back(std::move(lines[i]));
    if (!lines_ms[i].spans.empty()) minshape.rows.push_back(std::move(lines_ms[i]));
  }
}
// polygon.ml:608-609
inline void shapeminshape_of_unsorted_edgelist(std::vector<Edge> edges, Winding w, Shape& shape, Shape& minshape) {
  sort_edgelist_maxy_rev(edges);
  shapeminshape_of_edgelist(edges, w, false, shape, minshape);
}

// ---- Antialiasing tables, polygon.ml:616-671 ----
struct AATables {
  int maintable[aa_res][aa_res];          // [x][y]
  int volume;
  int table[aa_res][aa_res][aa_res];      // table[y][l-1][x] = 256 * sum of a run of length l starting at x in row y
  AATables() {
    auto pos = [](int p) { return ((double)(p - 1) * 6.) / (double)(aa_res - 1) - 3.; };
    auto sq = [](double x) { return x * x; };
    for (int x = 1; x <= aa_res; x++)
      for (int y = 1; y <= aa_res; y++)
        maintable[x - 1][y - 1] = (int)(std::exp(-((sq(pos(x)) + sq(pos(y))) / aa_softness)) * 255.);
    volume = gaussian(1, aa_res, 1, aa_res) / 255;
    for (int y = 0; y < aa_res; y++)
      for (int l = 1; l <= aa_res; l++)
        for (int x = 0; x < aa_res; x++)
          table[y][l - 1][x] = (x + l <= aa_res) ? gaussian(x + 1, x + l, y + 1, y + 1) : 0;
  }
  int gaussian(int x, int x2, int y, int y2) const {  // polygon.ml:636-643
    int t = 0;
    for (int xp = x - 1; xp <= x2 - 1; xp++)
      for (int yp = y - 1; yp <= y2 - 1; yp++) t += maintable[xp][yp];
    return t * 256;
  }
  int opacity_of_tableval(int t) const { return (t + volume / 2) / volume; }  // polygon.ml:650-651
  int lookup(int x, int y, int l) const { return table[y][l - 1][x]; }         // polygon.ml:670-671
};
inline const AATables& aa_tables() { static AATables t; return t; }

// polygon.ml:673-692
inline Shape mk_scaled_shape(Winding w, const std::vector<Edge>& edges) {
  Shape s, ms;
  if (edges.empty()) return s;
  int h = aa_res / 2;
  std::vector<Edge> scaled;
  for (const Edge& e : edges) scaled.push_back(Edge{e.x0 * h, e.x1 * h, e.y0 * h, e.y1 * h});
  shapeminshape_of_edgelist(scaled, w, true, s, ms);
  return s;
}
// polygon.ml:694-705 + sprite.ml:130-154 (shapespan_iter with clipping).
inline int pixel_coverage(const Shape& scaled, int x, int y) {
  const AATables& T = aa_tables();
  int h = aa_res / 2;
  int dx = -(x - 2) * h, dy = -(y - 2) * h;
  int minx = (x - 1) * h - h, miny = (y - 1) * h - h;
  int maxx = minx + aa_res - 1, maxy = miny + aa_res - 1;
  int count = 0;
  auto it = std::lower_bound(scaled.rows.begin(), scaled.rows.end(), miny,
                             [](const ShapeRow& r, int yy) { return r.y < yy; });
  for (; it != scaled.rows.end() && it->y <= maxy; ++it) {
    for (const Span& sp : it->spans) {
      int s = std::max(minx, sp.x), e = std::min(maxx, sp.x + sp.len - 1);
      int l = e - s + 1;
      if (l > 0) count += T.lookup(s + dx, it->y + dy, l);
    }
  }
  return count;
}
inline int pixel_opacity(const Shape& scaled, int x, int y) {
  return aa_tables().opacity_of_tableval(pixel_coverage(scaled, x, y));
}
}  // namespace oracle
