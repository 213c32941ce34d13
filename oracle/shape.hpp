// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's set-based span representation ("shapes") and its
// algebra.  Follows /root/reference/sprite.ml:23-54 (types), 201-239 (canonical form),
// 462-501 (box/translate), 542-549 (boxshape), 1180-1656 (union/difference/intersection),
// 1749-1877 (bloat/erode), 1973-1994 (point_in_shape).
// The reference stores a shape as bounds + vspans (maximal runs of consecutive rows);
// here a shape is the list of its non-empty rows in increasing y, each row a canonical
// span list (maximal x-runs, strictly separated).  The vspan grouping is a pure
// function of that list (consecutive y) and is re-derived on export (to_vspans).
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <algorithm>
#include <climits>
#include <stdexcept>
#include <vector>
namespace oracle {

struct Span { int x, len; };
inline bool operator==(const Span& a, const Span& b) { return a.x == b.x && a.len == b.len; }
typedef std::vector<Span> Spanline;
struct ShapeRow { int y; Spanline spans; };

struct Box { int x0 = 0, y0 = 0, x1 = 0, y1 = 0; };

struct Shape {
  std::vector<ShapeRow> rows;  // empty == NullShape
  bool null() const { return rows.empty(); }
};

// sprite.ml:1180-1215 — union of two canonical spanlines (abutting spans fuse).
inline Spanline spanline_union(const Spanline& a, const Spanline& b) {
  Spanline out;
  size_t i = 0, j = 0;
  auto push = [&](Span s) {
    if (!out.empty()) {
      Span& t = out.back();
      int te = t.x + t.len - 1;
      if (s.x <= te + 1) {  // overlap or abut → fuse
        int se = s.x + s.len - 1;
        if (se > te) t.len = se - t.x + 1;
        return;
      }
    }
    out.push_back(s);
  };
  while (i < a.size() || j < b.size()) {
    if (j >= b.size() || (i < a.size() && a[i].x <= b[j].x)) push(a[i++]);
    else push(b[j++]);
  }
  return out;
}

// sprite.ml:1298-1370 — a \ b on canonical spanlines.
inline Spanline spanline_difference(const Spanline& a, const Spanline& b) {
  Spanline out;
  size_t j = 0;
  for (const Span& s : a) {
    int cur = s.x, e = s.x + s.len - 1;
    while (j < b.size() && b[j].x + b[j].len - 1 < cur) j++;
    size_t k = j;
    while (k < b.size() && b[k].x <= e) {
      if (b[k].x > cur) out.push_back({cur, b[k].x - cur});
      cur = std::max(cur, b[k].x + b[k].len);
      k++;
    }
    if (cur <= e) out.push_back({cur, e - cur + 1});
  }
  return out;
}

// sprite.ml:1519-1580 — a ∩ b on canonical spanlines.
inline Spanline spanline_intersection(const Spanline& a, const Spanline& b) {
  Spanline out;
  size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    int ea = a[i].x + a[i].len - 1, eb = b[j].x + b[j].len - 1;
    int s = std::max(a[i].x, b[j].x), e = std::min(ea, eb);
    if (s <= e) out.push_back({s, e - s + 1});
    if (ea < eb) i++; else j++;
  }
  return out;
}

template <class F>
inline Shape rowwise(const Shape& a, const Shape& b, F f, bool keep_a_only, bool keep_b_only) {
  Shape out;
  size_t i = 0, j = 0;
  static const Spanline empty;
  while (i < a.rows.size() || j < b.rows.size()) {
    int ya = i < a.rows.size() ? a.rows[i].y : INT_MAX;
    int yb = j < b.rows.size() ? b.rows[j].y : INT_MAX;
    if (ya < yb) {
      if (keep_a_only) out.rows.push_back(a.rows[i]);
      i++;
    } else if (yb < ya) {
      if (keep_b_only) out.rows.push_back(b.rows[j]);
      j++;
    } else {
      Spanline r = f(a.rows[i].spans, b.rows[j].spans);
      if (!r.empty()) out.rows.push_back({ya, std::move(r)});
      i++; j++;
    }
  }
  return out;
}
inline Shape shape_union(const Shape& a, const Shape& b) {        // sprite.ml:1275-1293, `|||`
  return rowwise(a, b, spanline_union, true, true);
}
inline Shape shape_difference(const Shape& a, const Shape& b) {   // sprite.ml:1483-1512, `---`
  return rowwise(a, b, spanline_difference, true, false);
}
inline Shape shape_intersection(const Shape& a, const Shape& b) { // sprite.ml:1623-1656, `&&&`
  return rowwise(a, b, spanline_intersection, false, false);
}

// sprite.ml:462-465 — rectangle at raster resolution.
inline Shape shape_box(int x, int y, int w, int h) {
  Shape s;
  if (w == 0 && h == 0) return s;
  if (w < 0 || h < 0) throw std::runtime_error("Sprite.box: negative argument.");
  for (int r = 0; r < h; r++) s.rows.push_back({y + r, {{x, w}}});
  return s;
}
// sprite.ml:470-484
inline Shape translate_shape(int dx, int dy, const Shape& s) {
  Shape o = s;
  for (auto& r : o.rows) { r.y += dy; for (auto& sp : r.spans) sp.x += dx; }
  return o;
}
// sprite.ml:542-549 — tight bounds. Returns false for NullShape.
inline bool shape_bounds(const Shape& s, Box& b) {
  if (s.null()) return false;
  b.y0 = s.rows.front().y; b.y1 = s.rows.back().y;
  b.x0 = INT_MAX; b.x1 = INT_MIN;
  for (auto& r : s.rows) {
    b.x0 = std::min(b.x0, r.spans.front().x);
    b.x1 = std::max(b.x1, r.spans.back().x + r.spans.back().len - 1);
  }
  return true;
}
// sprite.ml:201-239 — canonical form: rows strictly increasing, no empty spanline,
// spans of positive length strictly separated (gap >= 1 pixel, i.e. not abutting).
inline bool shapecheck(const Shape& s) {
  for (size_t i = 0; i < s.rows.size(); i++) {
    if (i && s.rows[i].y <= s.rows[i - 1].y) return false;
    const Spanline& l = s.rows[i].spans;
    if (l.empty()) return false;
    for (size_t k = 0; k < l.size(); k++) {
      if (l[k].len <= 0) return false;
      if (k && l[k].x <= l[k - 1].x + l[k - 1].len) return false;
    }
  }
  return true;
}
inline long shape_card(const Shape& s) {  // sprite.ml:301-304
  long n = 0;
  for (auto& r : s.rows) for (auto& sp : r.spans) n += sp.len;
  return n;
}
// sprite.ml:1973-1994
inline bool point_in_shape(const Shape& s, int x, int y) {
  auto it = std::lower_bound(s.rows.begin(), s.rows.end(), y,
                             [](const ShapeRow& r, int yy) { return r.y < yy; });
  if (it == s.rows.end() || it->y != y) return false;
  for (auto& sp : it->spans) if (x >= sp.x && x < sp.x + sp.len) return true;
  return false;
}

// sprite.ml:1749-1864 — bloat m n = dilation by a (2m+1)x(2n+1) box.  The reference
// does an x-pass (`bloat_spanline`, 1754-1756: grow every span by m each side, fuse)
// then a y-pass of "rolling unions" (`bloatv`, 1838-1847: output row i = union of
// x-bloated rows i-n..i+n, rows outside the shape being empty).  The union tree is only
// a speed-up of that definition.
inline Shape bloat(int m, int n, const Shape& s) {
  if (s.null()) return s;
  std::vector<ShapeRow> xb;
  for (auto& r : s.rows) {
    Spanline l;
    for (auto& sp : r.spans) {
      Span g{sp.x - m, sp.len + 2 * m};
      if (!l.empty() && g.x <= l.back().x + l.back().len) {
        int e = std::max(l.back().x + l.back().len - 1, g.x + g.len - 1);
        l.back().len = e - l.back().x + 1;
      } else l.push_back(g);
    }
    xb.push_back({r.y, std::move(l)});
  }
  Shape out;
  int y0 = s.rows.front().y - n, y1 = s.rows.back().y + n;
  size_t lo = 0;
  for (int y = y0; y <= y1; y++) {
    while (lo < xb.size() && xb[lo].y < y - n) lo++;
    Spanline acc;
    for (size_t k = lo; k < xb.size() && xb[k].y <= y + n; k++) acc = spanline_union(acc, xb[k].spans);
    if (!acc.empty()) out.rows.push_back({y, std::move(acc)});
  }
  return out;
}
// sprite.ml:1867-1877
inline Shape erode(int m, int n, const Shape& s) {
  Box b;
  if (!shape_bounds(s, b)) return s;
  Shape enclosing = shape_box(b.x0 - m, b.y0 - n, b.x1 - b.x0 + 1 + 2 * m, b.y1 - b.y0 + 1 + 2 * n);
  return shape_difference(s, bloat(m, n, shape_difference(enclosing, s)));
}

// Flat interchange form used across the C API of both the oracle and the CUDA
// library: for every non-empty row, in increasing y: y, nspans, then nspans (x,len).
inline std::vector<int> shape_to_flat(const Shape& s) {
  std::vector<int> v;
  for (auto& r : s.rows) {
    v.push_back(r.y); v.push_back((int)r.spans.size());
    for (auto& sp : r.spans) { v.push_back(sp.x); v.push_back(sp.len); }
  }
  return v;
}
inline Shape shape_from_flat(const int* p, int n) {
  Shape s;
  int i = 0;
  while (i < n) {
    ShapeRow r; r.y = p[i++]; int k = p[i++];
    for (int q = 0; q < k; q++) { r.spans.push_back({p[i], p[i + 1]}); i += 2; }
    s.rows.push_back(std::move(r));
  }
  return s;
}
}  // namespace oracle
