// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's stroker.  Follows /root/reference/shapes.ml:13 (kappa), 17-23 (rotation),
// 28-30 (angle_to), 46-66 (quarter), 74-94 (joinsegs), 101-131 (arc), 134-143 (circle), 203-228 (mkcap), 246-252
// (reverserail), 257-288 (capsegment), 293-317 (crosspoint), 325-332 (point_possibly_on_lines), 337-350 (mkjoin),
// 359-421 (joinsegments), 425-434 (straight), 437-470 (bezier), 473-481 (strokesubpath), 484-487 (strokepath_inner),
// 499-516 (degenerate, clean_path), 519-530 (strokepath_polygon, strokepath).
// Third-party arithmetic not under /root/reference (camlpdf, unpinned in Makefile:20), restated from its published
// source: Pdfutil vector helpers (mkvector, invert, offset_point, perpendicular, scalevectolength, mkunitvector,
// between, distance_between, couple, pair_reduce, extremes) and Pdftransform (matrix_of_transform = fold of
// matrix_compose over [Translate; Scale; Rotate], transform_matrix).  With the origins at (0, 0) every entry of the
// composed matrix is one product (r * cos s, r * sin s) or a copy, so only transform_matrix's association order
// (x * a + y * c) + e matters.
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cmath>
#include <list>
#include "brush.hpp"   // Segment, Subpath, Path
namespace oracle {

enum Cap { ButtCap = 0, RoundCap = 1, ProjectingCap = 2 };
enum Join { RoundJoin = 0, MitredJoin = 1, BevelJoin = 2 };
struct StrokeSpec { Cap startcap; Join join; Cap endcap; double mitrelimit, linewidth; };
typedef std::vector<Segment> Rail;

const double shapes_pi = 4. * std::atan(1.);                       // Pdfutil.pi
const double kappa = ((std::sqrt(2.) - 1.) / 3.) * 4.;             // shapes.ml:13

// ---- Pdfutil vector helpers ----
inline Pt mkvector(Pt a, Pt b) { return Pt(b.first - a.first, b.second - a.second); }
inline Pt invert(Pt a) { return Pt(-a.first, -a.second); }
inline Pt offset_point(Pt v, Pt p) { return Pt(p.first + v.first, p.second + v.second); }
inline Pt perpendicular(Pt a) { return Pt(-a.second, a.first); }
inline Pt scalevectolength(double l, Pt v) {
  double cur = std::sqrt(v.first * v.first + v.second * v.second);
  if (cur == 0.) return v;
  double factor = l / cur;
  return Pt(v.first * factor, v.second * factor);
}
inline Pt mkunitvector(Pt s, Pt e) { return scalevectolength(1., mkvector(s, e)); }
inline Pt between(Pt a, Pt b) { return Pt((a.first + b.first) / 2., (a.second + b.second) / 2.); }
inline double distance_between(Pt p, Pt q) {
  auto sqr = [](double x) { return x * x; };
  return std::sqrt(sqr(p.first - q.first) + sqr(q.second - p.second));
}
inline Segment Straight(Pt a, Pt b) { Segment s; s.bezier = false; s.p[0] = a; s.p[1] = b; s.p[2] = s.p[3] = Pt(0., 0.); return s; }
inline Segment Bezier(Pt a, Pt b, Pt c, Pt d) { Segment s; s.bezier = true; s.p[0] = a; s.p[1] = b; s.p[2] = c; s.p[3] = d; return s; }
inline Rail couple_straight(const std::vector<Pt>& pts) {   // couple (fun x y -> Straight (x, y))
  Rail r;
  for (size_t i = 0; i + 1 < pts.size(); i++) r.push_back(Straight(pts[i], pts[i + 1]));
  return r;
}

// shapes.ml:17-23
inline double rotation(Pt c, Pt p, Pt q) {
  double px = p.first - c.first, py = p.second - c.second, qx = q.first - c.first, qy = q.second - c.second;
  double a = px * qy - py * qx, b = px * qx + py * qy;
  return std::atan2(a, b);
}
// shapes.ml:28-30
inline double angle_to(Pt c, Pt p) {
  double r = std::atan2(p.second - c.second, p.first - c.first);
  return r < 0. ? r + 2. * shapes_pi : r;
}
// shapes.ml:46-59 with Pdftransform.transform [Translate c; Scale ((0,0), r, r); Rotate ((0,0), s)]
inline Segment quarter(double s, Pt c, double r) {
  const double ma = std::cos(s) * r, mb = std::sin(s) * r, mc = -std::sin(s) * r, md = std::cos(s) * r;
  auto tr = [&](Pt p) { return Pt(p.first * ma + p.second * mc + c.first, p.first * mb + p.second * md + c.second); };
  return Bezier(tr(Pt(1., 0.)), tr(Pt(1., kappa)), tr(Pt(kappa, 1.)), tr(Pt(0., 1.)));
}
// shapes.ml:62-66
inline Segment quarter_anticlockwise(double s, Pt c, double r) {
  Segment q = quarter(s, c, r);
  return Bezier(q.p[3], q.p[2], q.p[1], q.p[0]);
}
// shapes.ml:74-81
inline Rail joinsegs(Rail segs) {
  for (size_t i = 0; i + 1 < segs.size(); i++) {
    ORACLE_ASSERT(segs[i].bezier && segs[i + 1].bezier, "Shapes.joinsegs: Segment not supported");
    segs[i + 1].p[0] = segs[i].p[3];
  }
  return segs;
}
// shapes.ml:85-94
inline Rail joinsegs_ends(Pt p1, Pt p2, Rail segments) {
  Rail segs = joinsegs(segments);
  if (segs.empty()) return segs;
  ORACLE_ASSERT(segs.front().bezier && segs.back().bezier, "Shapes.joinsegs_ends: Segment not supported");
  segs.front().p[0] = p1;
  segs.back().p[3] = p2;
  return segs;
}
// polygon.ml:129-141 (bezier_split), first part only
inline Segment bezier_split_left(double t, const Segment& q) {
  auto divide = [t](Pt a, Pt b) { double t1 = 1. - t; return Pt(t1 * a.first + t * b.first, t1 * a.second + t * b.second); };
  Pt l2 = divide(q.p[0], q.p[1]), h = divide(q.p[1], q.p[2]);
  Pt l3 = divide(l2, h), r3 = divide(q.p[2], q.p[3]);
  Pt r2 = divide(h, r3);
  Pt l4 = divide(l3, r2);
  return Bezier(q.p[0], l2, l3, l4);
}
// shapes.ml:101-131
inline Rail arc(Pt p1, Pt p2, Pt c) {
  const double ninety = shapes_pi / 2.;
  double angletogo = rotation(c, p1, p2), abs_angle = angle_to(c, p1), r = distance_between(p1, c);
  const bool clockwise = angletogo > 0.;
  const double ninety_abs = clockwise ? ninety : -ninety;
  Rail segments;
  angletogo = std::fabs(angletogo);
  while (angletogo > 0.) {
    if (angletogo >= ninety) {
      angletogo = angletogo - ninety;
      segments.push_back(clockwise ? quarter(abs_angle, c, r) : quarter_anticlockwise(abs_angle, c, r));
      abs_angle = std::fmod(abs_angle + ninety_abs, 2. * shapes_pi);
    } else {
      Segment q = clockwise ? quarter(abs_angle, c, r) : quarter_anticlockwise(abs_angle, c, r);
      double portion_needed = angletogo / ninety;
      segments.push_back(bezier_split_left(portion_needed, q));
      angletogo = 0.;
    }
  }
  return joinsegs_ends(p1, p2, segments);
}
// shapes.ml:134-143
inline Subpath circle(double x, double y, double r) {
  return joinsegs({quarter(0., Pt(x, y), r), quarter(shapes_pi / 2., Pt(x, y), r), quarter(shapes_pi, Pt(x, y), r), quarter(3. * shapes_pi / 2., Pt(x, y), r)});
}

// shapes.ml:203-228
inline Rail mkcap(Cap captype, Pt p1, Pt p2, double width, Pt perp_vector) {
  switch (captype) {
    case ButtCap: return {Straight(p1, p2)};
    case ProjectingCap: {
      Pt hv = scalevectolength(width / 2., perp_vector);
      Pt p = offset_point(hv, p1), q = offset_point(hv, p2);
      return couple_straight({p1, p, q, p2});
    }
    default: {
      double radius = width / 2.;
      Pt midpoint = between(p1, p2);
      Pt perpscaled = scalevectolength(radius, perp_vector);
      Pt centrearc = offset_point(perpscaled, midpoint);
      double control_length = radius * kappa;
      Pt lvector = scalevectolength(control_length, perp_vector);
      Pt cleftvector = scalevectolength(control_length, mkvector(p2, p1));
      Pt crightvector = scalevectolength(control_length, mkvector(p1, p2));
      Pt p1_up = offset_point(lvector, p1), p2_up = offset_point(lvector, p2);
      Pt c_left = offset_point(cleftvector, centrearc), c_right = offset_point(crightvector, centrearc);
      return {Bezier(p1, p1_up, c_left, centrearc), Bezier(centrearc, c_right, p2_up, p2)};
    }
  }
}
// shapes.ml:246-252
inline Rail reverserail(const Rail& r) {
  Rail o;
  for (size_t i = r.size(); i-- > 0;) {
    const Segment& s = r[i];
    o.push_back(s.bezier ? Bezier(s.p[3], s.p[2], s.p[1], s.p[0]) : Straight(s.p[1], s.p[0]));
  }
  return o;
}
// shapes.ml:257-288
inline Rail capsegment(const StrokeSpec& spec, const Rail& r, const Rail& r2) {
  ORACLE_ASSERT(!r.empty() && !r2.empty(), "Shapes.capsegment: empty rail");
  ORACLE_ASSERT(!r.front().bezier && !r.back().bezier && !r2.front().bezier && !r2.back().bezier, "Shapes.capsegment: malformed rail");
  Pt p1, p4, v, v2;
  if (r.size() == 1) { Pt s = r[0].p[0], e = r[0].p[1]; p1 = s; p4 = e; v = mkunitvector(e, s); v2 = mkunitvector(s, e); }
  else { Pt s = r.front().p[0], m = r.front().p[1], n = r.back().p[0], e = r.back().p[1]; p1 = s; p4 = e; v = mkunitvector(m, s); v2 = mkunitvector(n, e); }
  Pt p2 = r2.front().p[0], p3 = r2.back().p[1];
  Rail out = mkcap(spec.startcap, p1, p2, spec.linewidth, v);
  out.insert(out.end(), r2.begin(), r2.end());
  Rail ec = mkcap(spec.endcap, p3, p4, spec.linewidth, v2);
  out.insert(out.end(), ec.begin(), ec.end());
  Rail rr = reverserail(r);
  out.insert(out.end(), rr.begin(), rr.end());
  return out;
}
// shapes.ml:293-317
inline Pt crosspoint(Pt p, Pt v, Pt q, Pt w) {
  double px = p.first, py = p.second, vx = v.first, vy = v.second, qx = q.first, qy = q.second, wx = w.first, wy = w.second;
  if (vy == 0. && wx == 0.) return Pt(qx, py);
  if (vx == 0. && wy == 0.) return Pt(px, qy);
  if (wx == 0.) return Pt(qx, (vy / vx) * (qx - px) + py);
  if (vx == 0.) return Pt(px, (wy / wx) * (px - qx) + qy);
  if (wy == 0.) return Pt((qy - py) / (vy / vx) + px, qy);
  if (vy == 0.) return Pt((py - qy) / (wy / wx) + qx, py);
  double m = vy / vx, m2 = wy / wx;
  double c = py + (-px * m), c2 = qy + (-qx * m2);
  double pp = m / m2;
  double c3 = c2 * pp;
  double ycoeff = 1. - pp;
  double y = (c - c3) / ycoeff;
  double x = (c - y) / -m;
  return Pt(x, y);
}
inline Pt crosspoint_lines(Pt a, Pt b, Pt c, Pt d) { return crosspoint(a, mkvector(a, b), c, mkvector(c, d)); }
// shapes.ml:325-332 (Pdfutil.fmin / fmax)
inline bool point_possibly_on_lines(Pt a, Pt c, Pt a2, Pt c2, Pt p) {
  auto fmin_ = [](double x, double y) { return x < y ? x : y; };
  auto fmax_ = [](double x, double y) { return x > y ? x : y; };
  double min_x = fmin_(a.first, c.first), max_x = fmax_(a.first, c.first), min_y = fmin_(a.second, c.second), max_y = fmax_(a.second, c.second);
  double min_x2 = fmin_(a2.first, c2.first), max_x2 = fmax_(a2.first, c2.first), min_y2 = fmin_(a2.second, c2.second), max_y2 = fmax_(a2.second, c2.second);
  double x = p.first, y = p.second;
  return (x >= min_x && x <= max_x && y >= min_y && y <= max_y) || (x >= min_x2 && x <= max_x2 && y >= min_y2 && y <= max_y2);
}
// shapes.ml:337-350
inline Rail mkjoin(const StrokeSpec& spec, Join join, Pt c, Pt p1, Pt p2, Pt v1, Pt v2) {
  switch (join) {
    case BevelJoin: return {Straight(p1, p2)};
    case RoundJoin: return arc(p1, p2, c);
    default: {
      double angle_between = std::fabs(rotation(c, p1, p2)), phi = 2. * std::asin(1. / spec.mitrelimit);
      if (angle_between < phi) return mkjoin(spec, BevelJoin, c, p1, p2, v1, v2);
      Pt cp = crosspoint(p1, v1, p2, v2);
      return couple_straight({p1, cp, p2});
    }
  }
}
typedef std::pair<Rail, Rail> Rails;
// shapes.ml:359-421
inline Rails joinsegments(const StrokeSpec& spec, const Rails& L, const Rails& R) {
  const Rail &s1 = L.first, &s2 = L.second, &t1 = R.first, &t2 = R.second;
  ORACLE_ASSERT(!(s1.empty() && s2.empty()) && !(t1.empty() && t2.empty()), "Shapes.joinsegments: empty section");
  ORACLE_ASSERT(!s1.empty() && !s2.empty() && !t1.empty() && !t2.empty(), "Shapes.joinsegments: empty rail (hd / last of [])");
  const Segment ab = s1.back(), cd = s2.back(), a2b2 = t1.front(), c2d2 = t2.front();
  ORACLE_ASSERT(!ab.bezier && !cd.bezier && !a2b2.bezier && !c2d2.bezier, "joinsegments: Not implemented");
  Rail left1(s1.begin(), s1.end() - 1), left2(s2.begin(), s2.end() - 1), right1(t1.begin() + 1, t1.end()), right2(t2.begin() + 1, t2.end());
  Pt a = ab.p[0], b = ab.p[1], c = cd.p[0], d = cd.p[1], a2 = a2b2.p[0], b2 = a2b2.p[1], c2 = c2d2.p[0], d2 = c2d2.p[1];
  Pt lr_cross = crosspoint_lines(a, b, a2, b2), l2r2_cross = crosspoint_lines(c, d, c2, d2);
  bool on1 = point_possibly_on_lines(a, b, a2, b2, lr_cross), on2 = point_possibly_on_lines(c, d, c2, d2, l2r2_cross);
  auto cat = [](Rail x, const Rail& y, const Rail& z) { x.insert(x.end(), y.begin(), y.end()); x.insert(x.end(), z.begin(), z.end()); return x; };
  if (on1 == on2) return Rails(cat(left1, {Straight(a, b2)}, right1), cat(left2, {Straight(c, d2)}, right2));
  if (!on1) {   // join on rail lr
    Pt centre = between(b, d), vl = mkunitvector(a, b), vr = mkunitvector(b2, a2);
    Rail mid = {ab};
    Rail join = mkjoin(spec, spec.join, centre, b, a2, vl, vr);
    mid.insert(mid.end(), join.begin(), join.end());
    mid.push_back(a2b2);
    return Rails(cat(left1, mid, right1), cat(left2, {Straight(c, l2r2_cross), Straight(l2r2_cross, d2)}, right2));
  }
  Pt centre = between(b, d), vl = mkunitvector(c, d), vr = mkunitvector(d2, c2);
  Rail mid = {cd};
  Rail join = mkjoin(spec, spec.join, centre, d, c2, vl, vr);
  mid.insert(mid.end(), join.begin(), join.end());
  mid.push_back(c2d2);
  return Rails(cat(left1, {Straight(a, lr_cross), Straight(lr_cross, b2)}, right1), cat(left2, mid, right2));
}
// shapes.ml:425-434
inline Rails straight_rails(Pt s, Pt e, double width) {
  Pt offset = perpendicular(mkvector(s, e));
  Pt so = scalevectolength(width / 2., offset), so2 = invert(so);
  Pt a = offset_point(so, s), b = offset_point(so2, s), c = offset_point(so2, e), d = offset_point(so, e);
  return Rails({Straight(a, d)}, {Straight(b, c)});
}
// shapes.ml:437-470
inline Rails bezier_rails(Pt p1, Pt p2, Pt p3, Pt p4, double width) {
  std::vector<std::pair<Pt, Pt>> sub;
  bezier_subdivide(curve_accuracy, p1, p2, p3, p4, sub);
  std::vector<Pt> points;
  for (auto& e : sub) points.push_back(e.first);
  points.push_back(sub.back().second);
  std::vector<Pt> mid;   // midedge_offsets
  for (size_t i = 0; i + 1 < points.size(); i++) mid.push_back(perpendicular(mkvector(points[i], points[i + 1])));
  std::vector<Pt> offs;
  offs.push_back(mid.front());
  for (size_t i = 0; i + 1 < mid.size(); i++) offs.push_back(between(mid[i], mid[i + 1]));
  offs.push_back(mid.back());
  std::vector<Pt> op, op2;
  for (size_t i = 0; i < points.size(); i++) {
    Pt o = scalevectolength(width / 2., offs[i]);
    op.push_back(offset_point(o, points[i]));
    op2.push_back(offset_point(invert(o), points[i]));
  }
  return Rails(couple_straight(op), couple_straight(op2));
}
// Pdfutil.pair_reduce: apply f to neighbouring pairs (an odd one out is kept) until one element remains
inline Rails pair_reduce_join(const StrokeSpec& spec, std::vector<Rails> l) {
  ORACLE_ASSERT(!l.empty(), "pair_reduce: empty list");
  while (l.size() > 1) {
    std::vector<Rails> n;
    for (size_t i = 0; i + 1 < l.size(); i += 2) n.push_back(joinsegments(spec, l[i], l[i + 1]));
    if (l.size() & 1) n.push_back(l.back());
    l.swap(n);
  }
  return l[0];
}
// shapes.ml:473-481
inline Subpath strokesubpath(const StrokeSpec& spec, const Subpath& segments) {
  std::vector<Rails> rails;
  for (const Segment& s : segments)
    rails.push_back(s.bezier ? bezier_rails(s.p[0], s.p[1], s.p[2], s.p[3], spec.linewidth) : straight_rails(s.p[0], s.p[1], spec.linewidth));
  Rails j = pair_reduce_join(spec, rails);
  return capsegment(spec, j.first, j.second);
}
// shapes.ml:499-530: degenerate paths, clean_path, strokepath_polygon.  Returns the winding rule of the outline.
inline Winding strokepath_polygon(const StrokeSpec& spec, const Path& subpaths, Path& out) {
  out.clear();
  if (subpaths.size() == 1 && subpaths[0].size() == 1 && spec.startcap == RoundCap && spec.endcap == RoundCap) {
    const Segment& s = subpaths[0][0];
    bool deg = s.bezier ? (s.p[0] == s.p[1] && s.p[2] == s.p[3] && s.p[1] == s.p[2]) : (s.p[0] == s.p[1]);
    if (deg) { out.push_back(circle(s.p[0].first, s.p[0].second, spec.linewidth / 2.)); return NonZero; }
  }
  for (const Subpath& sp : subpaths) {
    Subpath clean;
    for (const Segment& s : sp) {
      bool ok = s.bezier ? !(s.p[0] == s.p[1] || s.p[2] == s.p[3]) : !(s.p[0] == s.p[1]);
      if (ok) clean.push_back(s);
    }
    if (!clean.empty()) out.push_back(strokesubpath(spec, clean));
  }
  return EvenOdd;
}
// shapes.ml:529-530 with polygon.ml:262-287: the sorted sub-bin edge list of the stroke's outline
inline std::vector<Edge> strokepath(const StrokeSpec& spec, const Path& path) {
  Path outline;
  strokepath_polygon(spec, path, outline);
  std::vector<Edge> es;
  for (const Subpath& sp : outline)
    for (const Segment& s : sp) {
      std::vector<std::pair<Pt, Pt>> e;
      if (!s.bezier) e.push_back({s.p[0], s.p[1]});
      else bezier_subdivide(curve_accuracy, s.p[0], s.p[1], s.p[2], s.p[3], e);
      for (auto& pe : e) es.push_back(Edge{sub_of_float(pe.first.first), sub_of_float(pe.second.first), sub_of_float(pe.first.second), sub_of_float(pe.second.second)});
    }
  sort_edgelist_maxy_rev(es);
  return es;
}

// polygon.ml:404-438 (bounds_polygon) and shapes.ml:522-540 (bounds_stroke): xmin, xmax, ymin, ymax
inline void bounds_polygon(const Path& subpaths, int b[4]) {
  ORACLE_ASSERT(!subpaths.empty(), "Polygon2.bounds_polygon: Malformed (empty) path");
  int minx = INT32_MAX, maxx = INT32_MIN, miny = INT32_MAX, maxy = INT32_MIN;   // (max_int / min_int of the reference: never met by data)
  for (const Subpath& sp : subpaths)
    for (const Segment& s : sp) {
      if (!s.bezier) {
        int x0 = pix_of_float(s.p[0].first), x1 = pix_of_float(s.p[1].first), y0 = pix_of_float(s.p[0].second), y1 = pix_of_float(s.p[1].second);
        minx = std::min(minx, std::min(x0, x1)); maxx = std::max(maxx, std::max(x0, x1));
        miny = std::min(miny, std::min(y0, y1)); maxy = std::max(maxy, std::max(y0, y1));
      } else {
        std::vector<std::pair<Pt, Pt>> e;
        bezier_subdivide(1., s.p[0], s.p[1], s.p[2], s.p[3], e);
        Subpath segs;
        for (auto& pe : e) segs.push_back(Straight(pe.first, pe.second));
        int bb[4];
        bounds_polygon(Path{segs}, bb);
        minx = std::min(minx, bb[0] - 1); maxx = std::max(maxx, bb[1] + 1);
        miny = std::min(miny, bb[2] - 1); maxy = std::max(maxy, bb[3] + 1);
      }
    }
  b[0] = minx; b[1] = maxx; b[2] = miny; b[3] = maxy;
}
inline void bounds_stroke(const Path& path, const StrokeSpec& spec, int b[4]) {
  double oversize = (spec.startcap == ProjectingCap || spec.endcap == ProjectingCap) ? spec.linewidth : spec.linewidth / 2.;
  double oversize2 = oversize;
  if (spec.join == MitredJoin) { double m = spec.mitrelimit * spec.linewidth; oversize2 = oversize > m ? oversize : m; }   // Pdfutil.fmax
  int oi = (int)std::ceil(oversize2);
  bounds_polygon(path, b);
  b[0] -= oi; b[1] += oi; b[2] -= oi; b[3] += oi;
}

}  // namespace oracle
