// host_geom.cpp — host-side geometry preparation (the step immediately before the raster path,
// SURVEY.md §8f N2): path flattening to integer sub-bin edges and brush-stroke sampling.  In the
// reference this is FP64 OCaml (polygon.ml:83-127, 143-218, 262-287; brush.ml:25-28, 126-130, 172;
// coord.ml:47) and stays on the host here as well; these entry points exist so that a host written
// in any language (and the Python scene builder for 10^5-object scenes) gets the same arithmetic.
// Plain C++ with one rounding per operation (-ffp-contract=off): no device work, no fallback role.
#include <math.h>
#include <stdint.h>
#include <utility>
#include <vector>
#include "../../include/coherence_b200.h"

namespace {
typedef std::pair<double, double> Pt;
const double kCurveAccuracy = 0.2;  // polygon.ml:19

double distance_point_from_line(Pt c, Pt a, Pt b) {  // polygon.ml:83-90
  double l = sqrt((b.first - a.first) * (b.first - a.first) + (b.second - a.second) * (b.second - a.second));
  double s = ((a.second - c.second) * (b.first - a.first) - (a.first - c.first) * (b.second - a.second)) / (l * l);
  return fabs(s) * l;
}
bool flat_enough(double eps, Pt p1, Pt p2, Pt p3, Pt p4) {  // polygon.ml:107-114
  double d1 = distance_point_from_line(p2, p1, p4), d2 = distance_point_from_line(p3, p1, p4);
  if (fpclassify(d1) == FP_NORMAL && fpclassify(d2) == FP_NORMAL) return d1 < eps && d2 < eps;
  return true;
}
void subdivide(double eps, Pt p1, Pt p2, Pt p3, Pt p4, std::vector<std::pair<Pt, Pt>>& out) {  // polygon.ml:119-127
  if (flat_enough(eps, p1, p2, p3, p4)) { out.push_back({p1, p4}); return; }
  auto half = [](Pt a, Pt b) { return Pt((a.first + b.first) / 2., (a.second + b.second) / 2.); };
  Pt l2 = half(p1, p2), h = half(p2, p3), l3 = half(l2, h), r3 = half(p3, p4), r2 = half(h, r3), l4 = half(l3, r2);
  subdivide(eps, p1, l2, l3, l4, out);
  subdivide(eps, l4, r2, r3, p4, out);
}
int sub_of_float(double f) { return (int)ceil(f * 32.0 - 16.0); }  // coord.ml:47
// segs: records of 9 doubles: kind (0 straight / 1 bezier) then up to 4 points
void flatten(const double* segs, int n, std::vector<std::pair<Pt, Pt>>& out) {
  for (int i = 0; i < n; i++) {
    const double* s = segs + 9 * i;
    if (s[0] == 0.) out.push_back({Pt(s[1], s[2]), Pt(s[3], s[4])});
    else subdivide(kCurveAccuracy, Pt(s[1], s[2]), Pt(s[3], s[4]), Pt(s[5], s[6]), Pt(s[7], s[8]), out);
  }
}
}  // namespace

extern "C" {
// Polygon.edgelist_of_path for one subpath: returns the number of edges; writes up to cap edges.
int64_t coh_host_edgelist_of_subpath(const double* segs, int32_t n_segs, int32_t* edges_out, int64_t cap) {
  std::vector<std::pair<Pt, Pt>> f;
  flatten(segs, n_segs, f);
  for (size_t i = 0; i < f.size() && (int64_t)i < cap; i++) {
    edges_out[4 * i] = sub_of_float(f[i].first.first); edges_out[4 * i + 1] = sub_of_float(f[i].first.second);
    edges_out[4 * i + 2] = sub_of_float(f[i].second.first); edges_out[4 * i + 3] = sub_of_float(f[i].second.second);
  }
  return (int64_t)f.size();
}
// Brush.points_of_brushstroke_smear (brush.ml:239-257) + the integer points of find_smear_directions (259-283).
namespace {
void subdivide_adjacent(Pt p1, Pt p2, Pt p3, Pt p4, std::vector<Pt>& starts) {
  const double d = sqrt((p1.first - p4.first) * (p1.first - p4.first) + (p1.second - p4.second) * (p1.second - p4.second));
  if (d <= 2.) { starts.push_back(p1); return; }
  auto half = [](Pt a, Pt b) { return Pt((a.first + b.first) / 2., (a.second + b.second) / 2.); };
  Pt l2 = half(p1, p2), h = half(p2, p3), l3 = half(l2, h), r3 = half(p3, p4), r2 = half(h, r3), l4 = half(l3, r2);
  subdivide_adjacent(p1, l2, l3, l4, starts);
  subdivide_adjacent(l4, r2, r3, p4, starts);
}
}  // namespace
int64_t coh_host_smear_points(const double* segs, int32_t n_segs, int32_t* points_out, int64_t cap) {
  std::vector<Pt> pts;
  for (int i = 0; i < n_segs; i++) {
    const double* s = segs + 9 * i;
    if (s[0] == 0.) {
      Pt a(s[1], s[2]), b(s[3], s[4]), m((a.first + b.first) / 2., (a.second + b.second) / 2.);   // Pdfutil.between
      subdivide_adjacent(a, m, m, b, pts);
    } else subdivide_adjacent(Pt(s[1], s[2]), Pt(s[3], s[4]), Pt(s[5], s[6]), Pt(s[7], s[8]), pts);
  }
  int64_t n = 0; int lx = 0, ly = 0;
  for (const Pt& p : pts) {
    const int x = (int)p.first, y = (int)p.second;   // toint: truncation
    if (n && x == lx && y == ly) continue;            // drop_duplicates
    if (n < cap) { points_out[2 * n] = x; points_out[2 * n + 1] = y; }
    lx = x; ly = y; n++;
  }
  return n;
}
// Brush.points_of_brushstroke rounded as in brush.ml:172, one subpath: returns the number of points.
int64_t coh_host_brush_points(const double* segs, int32_t n_segs, double radius, int32_t* points_out, int64_t cap) {
  const int w = (int)ceil(radius) * 2 + 1;       // brush.ml:25-28
  const double sep = (double)w / 20.;            // brush.ml:126-130
  std::vector<std::pair<Pt, Pt>> work;           // polygon.ml:186-204: flattened pieces of each segment are
  for (int i = 0; i < n_segs; i++) {             // prepended as a block (segment order reversed)
    std::vector<std::pair<Pt, Pt>> f;
    flatten(segs + 9 * i, 1, f);
    work.insert(work.begin(), f.begin(), f.end());
  }
  int64_t n = 0;
  size_t i = 0;
  while (i < work.size()) {                      // polygon.ml:173-184 takelength / 151-160 splitat
    double want = sep; bool found = false;
    while (i < work.size()) {
      Pt p1 = work[i].first, p2 = work[i].second;
      double l = sqrt((p2.first - p1.first) * (p2.first - p1.first) + (p2.second - p1.second) * (p2.second - p1.second));
      if (want <= l) {
        double prop = want / l;
        Pt p(p1.first * (1. - prop) + p2.first * prop, p1.second * (1. - prop) + p2.second * prop);
        if (n < cap) { points_out[2 * n] = (int)(p.first + 0.5); points_out[2 * n + 1] = (int)(p.second + 0.5); }
        n++;
        if (p == p2) i++; else work[i].first = p;
        found = true;
        break;
      }
      want -= l; i++;
    }
    if (!found) break;
  }
  return n;
}
}
