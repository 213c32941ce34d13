// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's integer convolution.  Follows
// /root/reference/convolve.ml:19-22 (kernels), 37-70 (mkunit, mkxy, gaussian),
// 115-158 (separable passes), 161-204 (running-sum box blur), 207-232 (dispatcher),
// 239-258 (convolve_sprite), 265-296 (convolve_sprite_in_shape) and
// /root/reference/sprite.ml:1887-1968 (horizontal / vertical span lists).
// FullKernel is not restated: it writes to (y,y) (convolve.ml:108) and no caller uses it.
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <cmath>
#include "sprite.hpp"
namespace oracle {

struct Kernel {
  enum Kind { Unit = 0, XY = 1 } kind = Unit;
  int radius = 1, total = 0;
  std::vector<int> values;  // XY only, 2r+1 taps
};
inline int radius_of_kernel(const Kernel& k) { return k.radius; }
inline Kernel mkunit(int r) {
  if (r <= 0) throw std::invalid_argument("Convolve.mkunit");
  Kernel k; k.kind = Kernel::Unit; k.radius = r; return k;
}
// convolve.ml:60-70: gaussian r x 0 = toint (4r^2 * (exp(-(x/r)^2) / 2) + 0.5)
inline Kernel mkgaussian(int r) {
  if (r <= 0) throw std::invalid_argument("Convolve.mkxy");
  Kernel k; k.kind = Kernel::XY; k.radius = r; k.total = 0;
  auto sq = [](double x) { return x * x; };
  for (int i = -r; i <= r; i++) {
    double g = std::exp(-(sq((double)i / (double)r) + sq(0. / (double)r))) / 2.;
    int v = (int)((double)(4 * r * r) * g + 0.5);
    k.values.push_back(v); k.total += v;
  }
  return k;
}
// convolve.ml:115-119 (quirk: the blue clamp is `min tb tb`)
inline colour setcanvas(int total, int tr, int tg, int tb, int ta) {
  tr /= total; tg /= total; tb /= total; ta /= total;
  tr = std::min(ta, tr); tg = std::min(ta, tg);
  return colour_of_rgba(tr, tg, tb, ta);
}
// convolve.ml:207-232 on a canvas pair; `shape` in canvas coordinates.
inline void convolve_canvas(Canvas& canvas, Canvas& canvas2, const Kernel& k, const Shape& shape) {
  int r = k.radius;
  // X pass: canvas -> canvas2 over the horizontal spans
  for (auto& row : shape.rows)
    for (auto& sp : row.spans) {
      if (k.kind == Kernel::XY) {
        for (int x = sp.x; x < sp.x + sp.len; x++) {
          int tr = 0, tg = 0, tb = 0, ta = 0;
          for (int q = -r; q <= r; q++) {
            int R, G, B, A; rgba_of_colour(canvas.at(x + q, row.y), R, G, B, A);
            int v = k.values[q + r];
            tr += R * v; tg += G * v; tb += B * v; ta += A * v;
          }
          canvas2.at(x, row.y) = setcanvas(k.total, tr, tg, tb, ta);
        }
      } else {
        int tr = 0, tg = 0, tb = 0, ta = 0, d = 2 * r + 1;
        for (int q = -r; q <= r; q++) { int R, G, B, A; rgba_of_colour(canvas.at(sp.x + q, row.y), R, G, B, A); tr += R; tg += G; tb += B; ta += A; }
        canvas2.at(sp.x, row.y) = colour_of_rgba(tr / d, tg / d, tb / d, ta / d);
        for (int x = sp.x + 1; x < sp.x + sp.len; x++) {
          int R, G, B, A; rgba_of_colour(canvas.at(x - r - 1, row.y), R, G, B, A); tr -= R; tg -= G; tb -= B; ta -= A;
          rgba_of_colour(canvas.at(x + r, row.y), R, G, B, A); tr += R; tg += G; tb += B; ta += A;
          canvas2.at(x, row.y) = colour_of_rgba(tr / d, tg / d, tb / d, ta / d);
        }
      }
    }
  // Y pass: canvas2 -> canvas over the vertical spans (maximal vertical runs per column,
  // sprite.ml:1899-1968).  Only the unit kernel's running sum depends on run starts.
  Box bb; if (!shape_bounds(shape, bb)) return;
  for (int x = bb.x0; x <= bb.x1; x++) {
    int runstart = 0; bool inrun = false;
    int tr = 0, tg = 0, tb = 0, ta = 0, d = 2 * r + 1;
    for (int y = bb.y0; y <= bb.y1 + 1; y++) {
      bool in = y <= bb.y1 && point_in_shape(shape, x, y);
      if (!in) { inrun = false; continue; }
      if (k.kind == Kernel::XY) {
        int sr = 0, sg = 0, sb = 0, sa = 0;
        for (int q = -r; q <= r; q++) {
          int R, G, B, A; rgba_of_colour(canvas2.at(x, y + q), R, G, B, A);
          int v = k.values[q + r];
          sr += R * v; sg += G * v; sb += B * v; sa += A * v;
        }
        canvas.at(x, y) = setcanvas(k.total, sr, sg, sb, sa);
      } else {
        if (!inrun) {
          runstart = y; (void)runstart; tr = tg = tb = ta = 0;
          for (int q = -r; q <= r; q++) { int R, G, B, A; rgba_of_colour(canvas2.at(x, y + q), R, G, B, A); tr += R; tg += G; tb += B; ta += A; }
        } else {
          int R, G, B, A; rgba_of_colour(canvas2.at(x, y - r - 1), R, G, B, A); tr -= R; tg -= G; tb -= B; ta -= A;
          rgba_of_colour(canvas2.at(x, y + r), R, G, B, A); tr += R; tg += G; tb += B; ta += A;
        }
        canvas.at(x, y) = colour_of_rgba(tr / d, tg / d, tb / d, ta / d);
      }
      inrun = true;
    }
  }
}
// convolve.ml:265-296 (and 239-258 with shape = pickup = bloat r r (shape of sprite)).
inline Sprite convolve_sprite_in_shape(const Kernel& k, const Sprite& spr, const Shape& shape, const Shape& pickup_shape) {
  Sprite none;
  if (spr.null()) return none;
  int r = k.radius;
  Box sb; shape_bounds(shape_of_sprite(spr), sb);
  Canvas canvas(sb.x0 - 2 * r, sb.y0 - 2 * r, sb.x1 - sb.x0 + 1 + 4 * r, sb.y1 - sb.y0 + 1 + 4 * r, clear_colour());
  flatten_sprite(spr, canvas);
  Canvas canvas2 = canvas;
  convolve_canvas(canvas, canvas2, k, shape);
  return pickup(pickup_shape, canvas);
}
inline Sprite convolve_sprite(const Kernel& k, const Sprite& spr) {
  if (spr.null()) return spr;
  Shape R = bloat(k.radius, k.radius, shape_of_sprite(spr));
  return convolve_sprite_in_shape(k, spr, R, R);
}
}  // namespace oracle
