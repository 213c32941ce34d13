// host_stroke.cpp — the stroker: Shapes.strokepath_polygon, shapes.ml:203-516 — the outline of a
// stroked path as closed subpaths of straight and bezier segments.  The flattening of that outline to sub-bin edges and
// its scan conversion run on the device (coh_strokepath / coh_shapeminshape_of_stroke in host_polygon.inl).
//
// Why the outline is built on the host: it is a short sequential fold per subpath (every join reads the segments the
// join before it rewrote), and mitred and round joins go through atan2 / asin / sin / cos, whose last bit decides where
// a sub-bin edge lands — the host's libm is the one the reference (OCaml's C runtime) calls; the device's is not.
//
// Layout: a rail is a std::list of segments, so that joining two stretches of a subpath is two pops and a splice; the
// joins are made in the order of Pdfutil.pair_reduce (neighbouring pairs, then pairs of pairs), which decides the
// rounding of the crossing points.  Plain C++, one rounding per operation (-ffp-contract=off).
#include <math.h>
#include <stdint.h>
#include <list>
#include <utility>
#include <vector>
#include "../../include/coherence_b200.h"

namespace {
struct V { double x, y; };
inline bool same(V a, V b) { return a.x == b.x && a.y == b.y; }
struct Seg { bool curve; V p[4]; };
typedef std::list<Seg> Rail;
struct RailPair { Rail l, r; };

const double kPi = 4. * atan(1.);
const double kKappa = ((sqrt(2.) - 1.) / 3.) * 4.;   // shapes.ml:13
const double kCurveAccuracy = 0.2;                   // polygon.ml:19

inline Seg line(V a, V b) { Seg s; s.curve = false; s.p[0] = a; s.p[1] = b; s.p[2] = s.p[3] = V{0., 0.}; return s; }
inline Seg curve(V a, V b, V c, V d) { Seg s; s.curve = true; s.p[0] = a; s.p[1] = b; s.p[2] = c; s.p[3] = d; return s; }
inline V sub(V a, V b) { return V{b.x - a.x, b.y - a.y}; }               // Pdfutil.mkvector a b
inline V add(V p, V v) { return V{p.x + v.x, p.y + v.y}; }               // Pdfutil.offset_point
inline V neg(V a) { return V{-a.x, -a.y}; }
inline V perp(V a) { return V{-a.y, a.x}; }
inline V mid(V a, V b) { return V{(a.x + b.x) / 2., (a.y + b.y) / 2.}; }
inline V to_length(double l, V v) {                                      // Pdfutil.scalevectolength
  const double len = sqrt(v.x * v.x + v.y * v.y);
  if (len == 0.) return v;
  const double f = l / len;
  return V{v.x * f, v.y * f};
}
inline V unit(V s, V e) { return to_length(1., sub(s, e)); }

// ---- flattening of a curve for its rails (polygon.ml:83-127) ----
double dist_from_chord(V c, V a, V b) {
  const double l = sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
  const double s = ((a.y - c.y) * (b.x - a.x) - (a.x - c.x) * (b.y - a.y)) / (l * l);
  return fabs(s) * l;
}
void chop(V p1, V p2, V p3, V p4, std::vector<V>& pts) {   // appends the start point of every piece
  const double d1 = dist_from_chord(p2, p1, p4), d2 = dist_from_chord(p3, p1, p4);
  const bool flat = (fpclassify(d1) == FP_NORMAL && fpclassify(d2) == FP_NORMAL) ? (d1 < kCurveAccuracy && d2 < kCurveAccuracy) : true;
  if (flat) { pts.push_back(p1); return; }
  const V l2 = mid(p1, p2), h = mid(p2, p3), l3 = mid(l2, h), r3 = mid(p3, p4), r2 = mid(h, r3), l4 = mid(l3, r2);
  chop(p1, l2, l3, l4, pts);
  chop(l4, r2, r3, p4, pts);
}

// ---- arcs (shapes.ml:17-30, 46-66, 101-131) ----
double turn(V c, V p, V q) {   // rotation
  const double px = p.x - c.x, py = p.y - c.y, qx = q.x - c.x, qy = q.y - c.y;
  return atan2(px * qy - py * qx, px * qx + py * qy);
}
Seg quarter(double s, V c, double r, bool anticlockwise) {
  // Pdftransform.transform [Translate c; Scale ((0, 0), r, r); Rotate ((0, 0), s)] of the standard quarter
  const double cs = cos(s), sn = sin(s);
  const double a = cs * r, b = sn * r, cc = -sn * r, d = cs * r;
  const V q[4] = {{1., 0.}, {1., kKappa}, {kKappa, 1.}, {0., 1.}};
  V t[4];
  for (int i = 0; i < 4; i++) t[i] = V{q[i].x * a + q[i].y * cc + c.x, q[i].x * b + q[i].y * d + c.y};
  return anticlockwise ? curve(t[3], t[2], t[1], t[0]) : curve(t[0], t[1], t[2], t[3]);
}
void arc(V p1, V p2, V c, Rail& out) {
  const double ninety = kPi / 2.;
  double togo = turn(c, p1, p2);
  double at = atan2(p1.y - c.y, p1.x - c.x);
  if (at < 0.) at = at + 2. * kPi;
  const double r = sqrt((p1.x - c.x) * (p1.x - c.x) + (c.y - p1.y) * (c.y - p1.y));
  const bool anti = !(togo > 0.);
  const double step = anti ? -ninety : ninety;
  togo = fabs(togo);
  Rail segs;
  while (togo > 0.) {
    Seg q = quarter(at, c, r, anti);
    if (togo >= ninety) {
      togo = togo - ninety;
      at = fmod(at + step, 2. * kPi);
    } else {   // the left part of Polygon.bezier_split (polygon.ml:129-141)
      const double t = togo / ninety, t1 = 1. - t;
      auto part = [t, t1](V a, V b) { return V{t1 * a.x + t * b.x, t1 * a.y + t * b.y}; };
      const V l2 = part(q.p[0], q.p[1]), h = part(q.p[1], q.p[2]), l3 = part(l2, h), r3 = part(q.p[2], q.p[3]), r2 = part(h, r3), l4 = part(l3, r2);
      q = curve(q.p[0], l2, l3, l4);
      togo = 0.;
    }
    if (!segs.empty()) q.p[0] = segs.back().p[3];   // joinsegs
    segs.push_back(q);
  }
  if (!segs.empty()) { segs.front().p[0] = p1; segs.back().p[3] = p2; }   // joinsegs_ends
  out.splice(out.end(), segs);
}

// ---- caps (shapes.ml:203-228) ----
void cap(int kind, V p1, V p2, double width, V out_dir, Rail& out) {
  if (kind == COH_CAP_BUTT) { out.push_back(line(p1, p2)); return; }
  if (kind == COH_CAP_PROJECTING) {
    const V hv = to_length(width / 2., out_dir), p = add(p1, hv), q = add(p2, hv);
    out.push_back(line(p1, p)); out.push_back(line(p, q)); out.push_back(line(q, p2));
    return;
  }
  const double radius = width / 2., control = radius * kKappa;
  const V top = add(mid(p1, p2), to_length(radius, out_dir));
  const V up = to_length(control, out_dir), left = to_length(control, sub(p2, p1)), right = to_length(control, sub(p1, p2));
  out.push_back(curve(p1, add(p1, up), add(top, left), top));
  out.push_back(curve(top, add(top, right), add(p2, up), p2));
}

// ---- joins (shapes.ml:293-421) ----
V cross(V p, V v, V q, V w) {   // crosspoint
  if (v.y == 0. && w.x == 0.) return V{q.x, p.y};
  if (v.x == 0. && w.y == 0.) return V{p.x, q.y};
  if (w.x == 0.) return V{q.x, (v.y / v.x) * (q.x - p.x) + p.y};
  if (v.x == 0.) return V{p.x, (w.y / w.x) * (p.x - q.x) + q.y};
  if (w.y == 0.) return V{(q.y - p.y) / (v.y / v.x) + p.x, q.y};
  if (v.y == 0.) return V{(p.y - q.y) / (w.y / w.x) + q.x, p.y};
  const double m = v.y / v.x, m2 = w.y / w.x;
  const double c = p.y + (-p.x * m), c2 = q.y + (-q.x * m2);
  const double ratio = m / m2;
  const double y = (c - c2 * ratio) / (1. - ratio);
  const double x = (c - y) / -m;
  return V{x, y};
}
bool in_box_of(V a, V b, V p) {
  const double x0 = a.x < b.x ? a.x : b.x, x1 = a.x > b.x ? a.x : b.x, y0 = a.y < b.y ? a.y : b.y, y1 = a.y > b.y ? a.y : b.y;
  return p.x >= x0 && p.x <= x1 && p.y >= y0 && p.y <= y1;
}
void join(const coh_strokespec& spec, V c, V p1, V p2, V v1, V v2, Rail& out) {   // mkjoin
  if (spec.join == COH_JOIN_ROUND) { arc(p1, p2, c, out); return; }
  if (spec.join == COH_JOIN_MITRED) {
    const double between = fabs(turn(c, p1, p2)), phi = 2. * asin(1. / spec.mitrelimit);
    if (!(between < phi)) {
      const V cp = cross(p1, v1, p2, v2);
      out.push_back(line(p1, cp)); out.push_back(line(cp, p2));
      return;
    }
  }
  out.push_back(line(p1, p2));
}
// joinsegments: A := A joined with B (B is emptied).  false: a rail ends in a curve (the reference fails).
bool join_rails(const coh_strokespec& spec, RailPair& A, RailPair& B) {
  if (A.l.empty() || A.r.empty() || B.l.empty() || B.r.empty()) return false;
  const Seg ab = A.l.back(), cd = A.r.back(), ab2 = B.l.front(), cd2 = B.r.front();
  if (ab.curve || cd.curve || ab2.curve || cd2.curve) return false;
  A.l.pop_back(); A.r.pop_back(); B.l.pop_front(); B.r.pop_front();
  const V a = ab.p[0], b = ab.p[1], c = cd.p[0], d = cd.p[1], a2 = ab2.p[0], b2 = ab2.p[1], c2 = cd2.p[0], d2 = cd2.p[1];
  const V xl = cross(a, sub(a, b), a2, sub(a2, b2)), xr = cross(c, sub(c, d), c2, sub(c2, d2));
  const bool on_l = in_box_of(a, b, xl) || in_box_of(a2, b2, xl), on_r = in_box_of(c, d, xr) || in_box_of(c2, d2, xr);
  if (on_l == on_r) {   // joined already, or the path goes back on itself
    A.l.push_back(line(a, b2)); A.r.push_back(line(c, d2));
  } else if (!on_l) {   // the outside of the corner is the l rail
    A.l.push_back(ab);
    join(spec, mid(b, d), b, a2, unit(a, b), unit(b2, a2), A.l);
    A.l.push_back(ab2);
    A.r.push_back(line(c, xr)); A.r.push_back(line(xr, d2));
  } else {
    A.r.push_back(cd);
    join(spec, mid(b, d), d, c2, unit(c, d), unit(d2, c2), A.r);
    A.r.push_back(cd2);
    A.l.push_back(line(a, xl)); A.l.push_back(line(xl, b2));
  }
  A.l.splice(A.l.end(), B.l); A.r.splice(A.r.end(), B.r);
  return true;
}

// rails of one segment (shapes.ml:425-470)
void rails_of(const Seg& s, double width, RailPair& R) {
  if (!s.curve) {
    const V o = to_length(width / 2., perp(sub(s.p[0], s.p[1]))), o2 = neg(o);
    R.l.push_back(line(add(s.p[0], o), add(s.p[1], o)));
    R.r.push_back(line(add(s.p[0], o2), add(s.p[1], o2)));
    return;
  }
  std::vector<V> pts;
  chop(s.p[0], s.p[1], s.p[2], s.p[3], pts);
  pts.push_back(s.p[3]);
  const size_t n = pts.size();   // >= 2
  std::vector<V> normal(n - 1);
  for (size_t i = 0; i + 1 < n; i++) normal[i] = perp(sub(pts[i], pts[i + 1]));
  V pl{0., 0.}, pr{0., 0.};
  for (size_t i = 0; i < n; i++) {
    const V raw = i == 0 ? normal[0] : (i == n - 1 ? normal[n - 2] : mid(normal[i - 1], normal[i]));
    const V o = to_length(width / 2., raw);
    const V l = add(pts[i], o), r = add(pts[i], neg(o));
    if (i) { R.l.push_back(line(pl, l)); R.r.push_back(line(pr, r)); }
    pl = l; pr = r;
  }
}

// strokesubpath (shapes.ml:473-481) + capsegment (257-288): the closed outline of one cleaned subpath
bool stroke_subpath(const coh_strokespec& spec, const std::vector<Seg>& segs, Rail& outline) {
  std::vector<RailPair> R(segs.size());
  for (size_t i = 0; i < segs.size(); i++) rails_of(segs[i], spec.linewidth, R[i]);
  // Pdfutil.pair_reduce: element i of a level joins element i + stride; the survivor of an odd level is carried along
  for (size_t stride = 1; stride < R.size(); stride *= 2)
    for (size_t i = 0; i + stride < R.size(); i += 2 * stride)
      if (!join_rails(spec, R[i], R[i + stride])) return false;
  Rail& l = R[0].l; Rail& r = R[0].r;
  if (l.empty() || r.empty() || l.front().curve || l.back().curve || r.front().curve || r.back().curve) return false;
  const V p1 = l.front().p[0], p4 = l.back().p[1], p2 = r.front().p[0], p3 = r.back().p[1];
  const V v_start = unit(l.front().p[1], l.front().p[0]), v_end = unit(l.back().p[0], l.back().p[1]);
  cap(spec.startcap, p1, p2, spec.linewidth, v_start, outline);
  outline.splice(outline.end(), r);
  cap(spec.endcap, p3, p4, spec.linewidth, v_end, outline);
  for (Rail::reverse_iterator it = l.rbegin(); it != l.rend(); ++it)   // reverserail
    outline.push_back(it->curve ? curve(it->p[3], it->p[2], it->p[1], it->p[0]) : line(it->p[1], it->p[0]));
  return true;
}
// Polygon.bounds_polygon (polygon.ml:404-438): pixel box of the segments' end points; a curve counts through its pieces at
// a flatness of 1, grown by one pixel.
inline int pix_of(double f) { const int sub = (int)ceil(f * 32.0 - 16.0); return (sub + 31) / 32; }   // coord.ml:44-50
struct PixBounds {
  int x0 = INT32_MAX, x1 = INT32_MIN, y0 = INT32_MAX, y1 = INT32_MIN;
  void point(V p, int grow) {
    const int x = pix_of(p.x), y = pix_of(p.y);
    if (x - grow < x0) x0 = x - grow;
    if (x + grow > x1) x1 = x + grow;
    if (y - grow < y0) y0 = y - grow;
    if (y + grow > y1) y1 = y + grow;
  }
};
void chop_coarse(V p1, V p2, V p3, V p4, PixBounds& b) {   // bezier_subdivide (bezier_epsilon 1.): both ends of every piece
  const double d1 = dist_from_chord(p2, p1, p4), d2 = dist_from_chord(p3, p1, p4);
  const bool flat = (fpclassify(d1) == FP_NORMAL && fpclassify(d2) == FP_NORMAL) ? (d1 < 1. && d2 < 1.) : true;
  if (flat) { b.point(p1, 1); b.point(p4, 1); return; }
  const V l2 = mid(p1, p2), h = mid(p2, p3), l3 = mid(l2, h), r3 = mid(p3, p4), r2 = mid(h, r3), l4 = mid(l3, r2);
  chop_coarse(p1, l2, l3, l4, b);
  chop_coarse(l4, r2, r3, p4, b);
}
}  // namespace

extern "C" {
// Shapes.strokepath_polygon (shapes.ml:484-526).  segs: 9-double records of all subpaths in order, subpath_segs[k] of them
// in subpath k.  Writes the outline's segments (min (n, cap_segs) records), the number of segments of each outline subpath
// (min (m, cap_subpaths) counts), m and the outline's winding rule (EvenOdd; NonZero for the circle of a degenerate path
// with round caps).  Returns n, or -1 where the reference fails (a rail that ends in a curve cannot be joined: only
// possible for malformed input).
int64_t coh_host_strokepath(const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                            double* segs_out, int64_t cap_segs, int32_t* subpath_segs_out, int32_t cap_subpaths,
                            int32_t* n_subpaths_out, int32_t* winding_out) {
  std::vector<Rail> outlines;
  int winding = COH_EVENODD;
  auto seg_at = [segs](int64_t i) {
    const double* s = segs + 9 * i;
    Seg g; g.curve = s[0] != 0.;
    for (int k = 0; k < 4; k++) g.p[k] = V{s[1 + 2 * k], s[2 + 2 * k]};
    return g;
  };
  bool circle = false;
  if (n_subpaths == 1 && subpath_segs[0] == 1 && spec->startcap == COH_CAP_ROUND && spec->endcap == COH_CAP_ROUND) {
    const Seg g = seg_at(0);   // shapes.ml:499-505, 519-523: a degenerate path with round caps is a circle
    if (g.curve ? (same(g.p[0], g.p[1]) && same(g.p[2], g.p[3]) && same(g.p[1], g.p[2])) : same(g.p[0], g.p[1])) {
      circle = true; winding = COH_NONZERO;
      Rail c;
      for (int q = 0; q < 4; q++) {
        Seg s = quarter(q == 0 ? 0. : (q == 1 ? kPi / 2. : (q == 2 ? kPi : 3. * kPi / 2.)), g.p[0], spec->linewidth / 2., false);
        if (q) s.p[0] = c.back().p[3];
        c.push_back(s);
      }
      outlines.push_back(std::move(c));
    }
  }
  if (!circle) {
    int64_t at = 0;
    for (int k = 0; k < n_subpaths; k++) {
      std::vector<Seg> clean;   // shapes.ml:507-516: zero-length lines and curves with a doubled end point are dropped
      for (int i = 0; i < subpath_segs[k]; i++, at++) {
        const Seg g = seg_at(at);
        if (g.curve ? !(same(g.p[0], g.p[1]) || same(g.p[2], g.p[3])) : !same(g.p[0], g.p[1])) clean.push_back(g);
      }
      if (clean.empty()) continue;
      Rail o;
      if (!stroke_subpath(*spec, clean, o)) return -1;
      outlines.push_back(std::move(o));
    }
  }
  int64_t n = 0; int32_t m = 0;
  for (const Rail& o : outlines) {
    if (m < cap_subpaths) subpath_segs_out[m] = (int32_t)o.size();
    m++;
    for (const Seg& s : o) {
      if (n < cap_segs) {
        double* d = segs_out + 9 * n;
        d[0] = s.curve ? 1. : 0.;
        for (int k = 0; k < 4; k++) { const bool used = s.curve || k < 2; d[1 + 2 * k] = used ? s.p[k].x : 0.; d[2 + 2 * k] = used ? s.p[k].y : 0.; }
      }
      n++;
    }
  }
  *n_subpaths_out = m; *winding_out = winding;
  return n;
}
// Shapes.bounds_stroke (shapes.ml:522-540): xmin, xmax, ymin, ymax of the path in pixels, grown by the stroke's reach.
// Returns -1 for a path without subpaths (the reference fails).
int32_t coh_host_bounds_stroke(const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths, int32_t bounds_out[4]) {
  if (n_subpaths <= 0) return -1;
  PixBounds b;
  int64_t at = 0;
  for (int k = 0; k < n_subpaths; k++)
    for (int i = 0; i < subpath_segs[k]; i++, at++) {
      const double* s = segs + 9 * at;
      if (s[0] == 0.) { b.point(V{s[1], s[2]}, 0); b.point(V{s[3], s[4]}, 0); }
      else chop_coarse(V{s[1], s[2]}, V{s[3], s[4]}, V{s[5], s[6]}, V{s[7], s[8]}, b);
    }
  double reach = (spec->startcap == COH_CAP_PROJECTING || spec->endcap == COH_CAP_PROJECTING) ? spec->linewidth : spec->linewidth / 2.;
  if (spec->join == COH_JOIN_MITRED) { const double m = spec->mitrelimit * spec->linewidth; if (!(reach > m)) reach = m; }
  const int grow = (int)ceil(reach);
  bounds_out[0] = b.x0 - grow; bounds_out[1] = b.x1 + grow; bounds_out[2] = b.y0 - grow; bounds_out[3] = b.y1 + grow;
  return 0;
}
}
