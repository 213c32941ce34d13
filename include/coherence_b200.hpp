// coherence_b200.hpp — C++ host side above the C ABI (coherence_b200.h), mirroring the module
// interfaces of the OCaml reference for the raster hot path: same names, argument meaning and error
// behaviour, so that code written against render.mli / polygon.mli / sprite.mli / colour.mli /
// cache.mli / convolve.mli / filters.mli reads the same here.  The OCaml side proper is
// ocaml/coherence_gpu.ml + ocaml/coherence_stubs.c (no OCaml toolchain in this image: source only);
// this header is the compiled-language mirror that IS built and tested (tests/cpp).
//
//   reference                                   here
//   Failure msg (failwith, 98 sites)            coherence::Failure (what() = the message)
//   Colour.colour (31-bit int, colour.ml:66)    Colour::colour = premultiplied RGBA8 word; codec in Colour::
//   Sprite.shape (sprite.ml:46-54)              Sprite::shape  (device-resident span set, RAII)
//   Polygon.edge {x0;y0;x1;y1} (polygon.ml:19)  Polygon::edge
//   Render.renderobject = Obj (idset, geom,     Render::renderobject (geometry tree; the affine transform is
//     transform, compop) (render.ml:19-75)        applied before this boundary)
//   Render.render_frame / render_simple_scene   Render::render_frame / render_simple_scene (into the device
//     (render.mli:211-217)                        framebuffer; read back with Render::read_rgba / read_rgb888)
//
// Header-only; link with libcoherence_b200.so.  Not thread-safe, like its single-threaded model.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "coherence_b200.h"

namespace coherence {

struct Failure : std::runtime_error {  // OCaml's Failure
  explicit Failure(const std::string& m) : std::runtime_error(m) {}
};

// The device context (one GPU, one scanline band).  The reference has global state (the cache, the
// canvas); here it hangs off one process-wide context created on first use.
class Context {
 public:
  static Context& get(int device = -1) {
    static Context c(device);
    return c;
  }
  coh_ctx* raw() const { return ctx_; }
  void check(int rc) const {
    if (rc) throw Failure(coh_last_error(ctx_));
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  ~Context() { coh_shutdown(ctx_); }

 private:
  explicit Context(int device) {
    if (coh_init(device, &ctx_)) throw Failure(coh_last_error(nullptr));
  }
  coh_ctx* ctx_ = nullptr;
};
inline void ck(int rc) { Context::get().check(rc); }

// ---- colour.mli ------------------------------------------------------------------------------
namespace Colour {
typedef uint32_t colour;  // r | g << 8 | b << 16 | a << 24, premultiplied
inline colour colour_of_rgba(int r, int g, int b, int a) {  // colour.ml:99-130 (asserts r, g, b <= a)
  if (r < 0 || g < 0 || b < 0 || a < 0 || a > 255 || r > a || g > a || b > a) throw Failure("Colour.colour_of_rgba");
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);
}
inline colour colour_of_rgba_float(double r, double g, double b, double a) {  // colour.ml:246-252: toint (x *. 255.)
  int ai = (int)(a * 255.);
  auto ch = [&](double v) { int c = (int)(v * 255.); return c < ai ? c : ai; };
  return colour_of_rgba(ch(r * a), ch(g * a), ch(b * a), ai);
}
inline int alpha_of_colour(colour c) { return (int)(c >> 24); }
inline int32_t ocaml_colour(colour c) { return coh_colour_of_rgba8(c); }      // the reference's 31-bit encoding
inline colour of_ocaml_colour(int32_t c) { return coh_rgba8_of_colour(c); }
inline int div255(int i) { return (i + (i >> 8) + 1) >> 8; }                   // colour.ml:287
inline colour dissolve(colour c, int delta) {                                  // colour.ml:291-304
  if (delta == 0) return 0;
  if (delta == 255) return c;
  return (uint32_t)div255((c & 255) * delta) | ((uint32_t)div255(((c >> 8) & 255) * delta) << 8) |
         ((uint32_t)div255(((c >> 16) & 255) * delta) << 16) | ((uint32_t)div255((c >> 24) * delta) << 24);
}
const colour clear = 0u, white = 0xFFFFFFFFu, black = 0xFF000000u, lightgrey = colour_of_rgba(211, 211, 211, 255);
}  // namespace Colour

// ---- sprite.mli (shapes) ------------------------------------------------------------------------
namespace Sprite {
class shape {  // NullShape = no handle
 public:
  shape() {}
  explicit shape(coh_shape_t h) : h_(h ? new Handle(h) : nullptr) {}
  coh_shape_t handle() const { return h_ ? h_->h : 0; }
  bool is_null() const { return !h_; }
  // canonical flat form: for every non-empty row in increasing y:  y, nspans, then (x, len) pairs
  std::vector<int32_t> spans() const {
    int64_t n = 0;
    ck(coh_shape_export_size(Context::get().raw(), handle(), &n));
    std::vector<int32_t> out((size_t)n);
    int64_t got = 0;
    if (n) ck(coh_shape_export(Context::get().raw(), handle(), out.data(), n, &got));
    return out;
  }

 private:
  struct Handle {
    coh_shape_t h;
    explicit Handle(coh_shape_t x) : h(x) {}
    ~Handle() { coh_shape_free(Context::get().raw(), h); }
  };
  std::shared_ptr<Handle> h_;
};
inline shape box(int x, int y, int w, int h) {  // sprite.ml:462 (Failure on negative extents)
  coh_shape_t o = 0;
  ck(coh_shape_box(Context::get().raw(), x, y, w, h, &o));
  return shape(o);
}
inline shape of_spans(const std::vector<int32_t>& flat) {
  coh_shape_t o = 0;
  ck(coh_shape_import(Context::get().raw(), flat.data(), (int64_t)flat.size(), &o));
  return shape(o);
}
#define COH_SHAPE_BINOP(name, fn)                                        \
  inline shape name(const shape& a, const shape& b) {                    \
    coh_shape_t o = 0;                                                   \
    ck(fn(Context::get().raw(), a.handle(), b.handle(), &o));            \
    return shape(o);                                                     \
  }
COH_SHAPE_BINOP(shape_union, coh_shape_union)                // ( ||| ) sprite.ml:1275
COH_SHAPE_BINOP(shape_difference, coh_shape_difference)      // ( --- ) sprite.ml:1483
COH_SHAPE_BINOP(shape_intersection, coh_shape_intersection)  // ( &&& ) sprite.ml:1623
#undef COH_SHAPE_BINOP
inline shape operator|(const shape& a, const shape& b) { return shape_union(a, b); }
inline shape operator-(const shape& a, const shape& b) { return shape_difference(a, b); }
inline shape operator&(const shape& a, const shape& b) { return shape_intersection(a, b); }
inline shape translate_shape(int dx, int dy, const shape& s) {  // sprite.ml:476
  coh_shape_t o = 0;
  ck(coh_shape_translate(Context::get().raw(), s.handle(), dx, dy, &o));
  return shape(o);
}
inline shape bloat(int m, int n, const shape& s) {  // sprite.ml:1857
  coh_shape_t o = 0;
  ck(coh_shape_bloat(Context::get().raw(), s.handle(), m, n, &o));
  return shape(o);
}
inline shape erode(int m, int n, const shape& s) {  // sprite.ml:1867
  coh_shape_t o = 0;
  ck(coh_shape_erode(Context::get().raw(), s.handle(), m, n, &o));
  return shape(o);
}
inline int64_t shape_card(const shape& s) {  // sprite.ml:301
  int64_t n = 0;
  ck(coh_shape_card(Context::get().raw(), s.handle(), &n));
  return n;
}
}  // namespace Sprite

// ---- fill.mli ------------------------------------------------------------------------------------
namespace Fill {
struct fill {  // the closure record of fill.ml:19-58 as a descriptor (closures cannot cross the ABI)
  int kind = COH_FILL_PLAIN, flags = 0;
  Colour::colour c0 = 0, c1 = 0;
  double p[6] = {0, 0, 0, 0, 0, 0};
};
inline fill plain(Colour::colour c) {  // fill.ml:62
  fill f; f.c0 = c; return f;
}
inline fill gradient(double x0, double y0, double x1, double y1, bool ext_s, bool ext_e, Colour::colour cs, Colour::colour ce) {  // fill.ml:77
  fill f; f.kind = COH_FILL_AXIAL; f.c0 = cs; f.c1 = ce; f.flags = (ext_s ? COH_FILL_EXT_S : 0) | (ext_e ? COH_FILL_EXT_E : 0);
  f.p[0] = x0; f.p[1] = y0; f.p[2] = x1; f.p[3] = y1; return f;
}
inline fill radial(double cx, double cy, double px, double py, double qx, double qy, bool ext_s, bool ext_e, Colour::colour cs, Colour::colour ce) {  // fill.ml:112
  fill f; f.kind = COH_FILL_RADIAL; f.c0 = cs; f.c1 = ce; f.flags = (ext_s ? COH_FILL_EXT_S : 0) | (ext_e ? COH_FILL_EXT_E : 0);
  f.p[0] = cx; f.p[1] = cy; f.p[2] = px; f.p[3] = py; f.p[4] = qx; f.p[5] = qy; return f;
}
}  // namespace Fill

// ---- polygon.mli -----------------------------------------------------------------------------------
namespace Polygon {
struct edge { int32_t x0, y0, x1, y1; };  // sub-pixel bins (polygon.ml:19-26)
enum winding_rule { NonZero = COH_NONZERO, EvenOdd = COH_EVENODD };
struct segment {  // Pdfgraphics.Straight / Bezier
  bool bezier; double p[8];
  static segment straight(double ax, double ay, double bx, double by) { return segment{false, {ax, ay, bx, by, 0, 0, 0, 0}}; }
  static segment curve(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy) { return segment{true, {ax, ay, bx, by, cx, cy, dx, dy}}; }
};
typedef std::vector<segment> subpath;
// Polygon.edgelist_of_path for one subpath (polygon.ml:119-127, 262-287)
inline std::vector<edge> edgelist_of_subpath(const subpath& sp) {
  std::vector<double> rec(9 * sp.size());
  for (size_t i = 0; i < sp.size(); i++) {
    rec[9 * i] = sp[i].bezier ? 1. : 0.;
    for (int k = 0; k < 8; k++) rec[9 * i + 1 + k] = sp[i].p[k];
  }
  int64_t n = coh_host_edgelist_of_subpath(rec.data(), (int32_t)sp.size(), nullptr, 0);
  std::vector<edge> out((size_t)n);
  if (n) coh_host_edgelist_of_subpath(rec.data(), (int32_t)sp.size(), (int32_t*)out.data(), n);
  return out;
}
inline subpath path_of_pointlist(const std::vector<std::pair<double, double>>& pts) {  // polygon.ml:66-76 (closed)
  subpath sp;
  for (size_t i = 0; i < pts.size(); i++) {
    const auto &a = pts[i], &b = pts[(i + 1) % pts.size()];
    sp.push_back(segment::straight(a.first, a.second, b.first, b.second));
  }
  return sp;
}
// Polygon.shapeminshape_of_unsorted_edgelist (polygon.ml:608-609)
inline std::pair<Sprite::shape, Sprite::shape> shapeminshape_of_unsorted_edgelist(const std::vector<edge>& edges, winding_rule w) {
  coh_shape_t s = 0, m = 0;
  ck(coh_shapeminshape_of_edgelist(Context::get().raw(), (const int32_t*)edges.data(), (int32_t)edges.size(), (int32_t)w, &s, &m));
  return {Sprite::shape(s), Sprite::shape(m)};
}
// AA opacity (0..255) of every pixel of `shp`, in span order (polygon.ml:616-746 pixel_coverage)
inline std::vector<uint8_t> polygon_opacity(const std::vector<edge>& edges, winding_rule w, const Sprite::shape& shp) {
  std::vector<uint8_t> out((size_t)Sprite::shape_card(shp));
  int64_t n = 0;
  ck(coh_polygon_opacity(Context::get().raw(), (const int32_t*)edges.data(), (int32_t)edges.size(), (int32_t)w, shp.handle(), out.data(), (int64_t)out.size(), &n));
  out.resize((size_t)n);
  return out;
}
}  // namespace Polygon

// ---- convolve.mli ------------------------------------------------------------------------------------
namespace Convolve {
struct kernel { int kind, radius; };
inline kernel mkunit(int r) {      // convolve.ml:37-44 (Invalid_argument on r <= 0)
  if (r <= 0) throw std::invalid_argument("Convolve.mkunit");
  return kernel{COH_CONV_UNIT, r};
}
inline kernel mkgaussian(int r) {  // convolve.ml:60-70
  if (r <= 0) throw std::invalid_argument("Convolve.mkxy");
  return kernel{COH_CONV_GAUSSIAN, r};
}
inline int radius_of_kernel(const kernel& k) { return k.radius; }
}  // namespace Convolve

// ---- cache.mli ------------------------------------------------------------------------------------
namespace Cache {
inline void usecache(bool on) { ck(coh_cache_configure(Context::get().raw(), on ? 1 : 0, 0)); }     // cache.mli:32
inline void setsize(int64_t bytes) { ck(coh_cache_configure(Context::get().raw(), 1, bytes)); }       // cache.mli:27-28
inline void clear() { ck(coh_cache_clear(Context::get().raw())); }
}  // namespace Cache

// ---- render.mli ------------------------------------------------------------------------------------
namespace Render {
typedef int64_t idset;  // Id.idset; negative = Id.new_ids () on every render (never cached)
struct compop { int pretrans = -1; };  // Over | PreTrans (v, Over) with toint (v *. 255.) (render.ml:1295-1298)
inline compop Over() { return compop{}; }
inline compop PreTrans(double v) { compop c; c.pretrans = (int)(v * 255.); return c; }

struct renderobject;
typedef std::vector<renderobject> scene;  // head = front-most
enum cpgop { Union = COH_CPG_UNION, Intersection = COH_CPG_INTERSECTION, Subtraction = COH_CPG_SUBTRACTION, ExclusiveOr = COH_CPG_EXCLUSIVEOR };
enum filterkind { Hole = COH_FILTER_HOLE, Monochrome = COH_FILTER_MONOCHROME, Blur = COH_FILTER_BLUR, ReadingScene = COH_FILTER_SCENE, Minus = COH_FILTER_MINUS };

// Obj (idset, geometry, transform, compop) with the transform already applied (render.ml:19-75)
struct renderobject {
  enum Geom { GPath, GStrokedPath, GCPG, GBrushstroke, GPrimitive, GGroup, GFilter } geom = GPath;
  idset id = -1;
  compop op;
  int dx = 0, dy = 0;  // translate_renderobject (render.ml:259-271): an alias of the cached geometry
  Fill::fill fill;
  std::vector<Polygon::edge> edges, edges_b;  // Path / StrokedPath outline / CPG operand a; CPG operand b
  Polygon::winding_rule winding = Polygon::NonZero, winding_b = Polygon::NonZero;
  cpgop cpg = Union;
  Convolve::kernel convolved{0, 0};           // Convolved (kernel, this geometry) when kind != 0
  double brush_opacity = 1., brush_radius = 1.;
  std::vector<std::pair<int, int>> stamps;    // Brush.points_of_brushstroke
  int prim[4] = {0, 0, 0, 0}; bool prim_null = false;  // Primitive (colour, Rectangle ...)
  scene members;                              // Group
  filterkind filter = Hole; Convolve::kernel filter_kernel{0, 0};
  std::shared_ptr<scene> reading_scene;       // Filter: the rewritten scene below (affine, rgb, wireframe ...)
};
inline renderobject Basic_Path(const Fill::fill& f, const std::vector<Polygon::subpath>& path, Polygon::winding_rule w = Polygon::NonZero, compop op = Over(), idset id = -1) {
  renderobject o; o.geom = renderobject::GPath; o.fill = f; o.winding = w; o.op = op; o.id = id;
  for (const auto& sp : path) { auto e = Polygon::edgelist_of_subpath(sp); o.edges.insert(o.edges.end(), e.begin(), e.end()); }
  return o;
}
inline renderobject Basic_CPG(const Fill::fill& f, cpgop c, const std::vector<Polygon::subpath>& a, const std::vector<Polygon::subpath>& b, compop op = Over(), idset id = -1) {
  renderobject o = Basic_Path(f, a, Polygon::NonZero, op, id);
  o.geom = renderobject::GCPG; o.cpg = c;
  for (const auto& sp : b) { auto e = Polygon::edgelist_of_subpath(sp); o.edges_b.insert(o.edges_b.end(), e.begin(), e.end()); }
  return o;
}
inline renderobject Convolved(const Convolve::kernel& k, renderobject basic) { basic.convolved = k; return basic; }  // render.ml:63
inline renderobject Primitive_Rectangle(Colour::colour c, double xmin, double ymin, double xmax, double ymax) {    // render.ml:573-586
  renderobject o; o.geom = renderobject::GPrimitive; o.fill = Fill::plain(c);
  o.prim[0] = (int)xmin; o.prim[1] = (int)ymin; o.prim[2] = (int)xmax; o.prim[3] = (int)ymax;
  return o;
}
inline renderobject Group(const scene& members, compop op = Over(), idset id = -1) {
  if (members.empty()) throw Failure("Empty groups aren't allowed");  // render.ml:317
  renderobject o; o.geom = renderobject::GGroup; o.members = members; o.op = op; o.id = id;
  return o;
}
inline renderobject Filter(filterkind k, const std::vector<Polygon::subpath>& geometry, const Fill::fill& matte = Fill::plain(Colour::white),
                           Convolve::kernel kern = Convolve::kernel{0, 0}, std::shared_ptr<scene> reading = nullptr) {
  renderobject o = Basic_Path(matte, geometry);
  o.geom = renderobject::GFilter; o.filter = k; o.filter_kernel = kern; o.reading_scene = reading;
  return o;
}
inline renderobject translate_renderobject(int dx, int dy, renderobject o) {  // render.ml:259-271
  if (o.geom == renderobject::GGroup) { for (auto& m : o.members) m = translate_renderobject(dx, dy, m); }
  else { o.dx += dx; o.dy += dy; }
  return o;
}

// Flattening of a scene list (+ background list) into the ABI's arrays.
class Flattened {
 public:
  std::vector<coh_object> objs;
  std::vector<int32_t> edges, points;
  int n_background = 0;
  void add_scene(const scene& s) {
    std::vector<std::pair<size_t, std::shared_ptr<scene>>> reading;
    for (const auto& o : s) add(o, reading);
    for (auto& r : reading) {  // reading-scene groups follow every ordinary scene object
      coh_object g = blank(COH_OBJ_GROUP_BEGIN); g.filter_kind = COH_FILTER_READING_SCENE;
      objs[r.first].first2 = (int32_t)objs.size();
      objs.push_back(g);
      std::vector<std::pair<size_t, std::shared_ptr<scene>>> none;
      for (const auto& o : *r.second) add(o, none);
      if (!none.empty()) throw Failure("filters inside a reading scene are not supported");
      objs.push_back(blank(COH_OBJ_GROUP_END));
    }
  }
  void add_background(const scene& s) {
    size_t before = objs.size();
    std::vector<std::pair<size_t, std::shared_ptr<scene>>> none;
    for (const auto& o : s) add(o, none);
    if (!none.empty()) throw Failure("filters in the background list are not supported");
    n_background += (int)(objs.size() - before);
  }

 private:
  static coh_object blank(int kind) {
    coh_object c = coh_object();
    c.kind = kind; c.pretrans = -1; c.id = -1;
    return c;
  }
  int32_t put_edges(const std::vector<Polygon::edge>& e) {
    int32_t first = (int32_t)(edges.size() / 4);
    for (const auto& x : e) { edges.push_back(x.x0); edges.push_back(x.y0); edges.push_back(x.x1); edges.push_back(x.y1); }
    return first;
  }
  void add(const renderobject& o, std::vector<std::pair<size_t, std::shared_ptr<scene>>>& reading) {
    coh_object c = blank(COH_OBJ_PATH);
    c.pretrans = o.op.pretrans; c.id = o.id; c.dx = o.dx; c.dy = o.dy;
    c.fill_kind = o.fill.kind; c.colour0 = o.fill.c0; c.colour1 = o.fill.c1; c.fill_flags = o.fill.flags;
    for (int k = 0; k < 6; k++) c.fparam[k] = o.fill.p[k];
    switch (o.geom) {
      case renderobject::GGroup:
        c.kind = COH_OBJ_GROUP_BEGIN; objs.push_back(c);
        for (const auto& m : o.members) add(m, reading);
        objs.push_back(blank(COH_OBJ_GROUP_END));
        return;
      case renderobject::GPrimitive:
        c.kind = COH_OBJ_PRIMITIVE; c.prim_null = o.prim_null ? 1 : 0;
        for (int k = 0; k < 4; k++) c.prim[k] = o.prim[k];
        break;
      case renderobject::GBrushstroke:
        c.kind = COH_OBJ_BRUSH; c.first = (int32_t)(points.size() / 2); c.count = (int32_t)o.stamps.size();
        for (const auto& p : o.stamps) { points.push_back(p.first); points.push_back(p.second); }
        c.brush_opacity = o.brush_opacity; c.brush_radius = o.brush_radius;
        break;
      case renderobject::GCPG:
        c.kind = COH_OBJ_CPG; c.first = put_edges(o.edges); c.count = (int32_t)o.edges.size(); c.winding = o.winding;
        c.first2 = put_edges(o.edges_b); c.count2 = (int32_t)o.edges_b.size(); c.winding2 = o.winding_b; c.cpg_op = o.cpg;
        break;
      case renderobject::GFilter:
        c.kind = COH_OBJ_FILTER; c.first = put_edges(o.edges); c.count = (int32_t)o.edges.size(); c.winding = o.winding;
        c.filter_kind = o.filter; c.filter_kernel = o.filter_kernel.kind | (o.filter_kernel.radius << 8);
        if (o.filter == ReadingScene) {
          if (!o.reading_scene) throw Failure("Filter: no reading scene");
          reading.push_back({objs.size(), o.reading_scene});
        }
        break;
      default:  // Path, StrokedPath
        c.first = put_edges(o.edges); c.count = (int32_t)o.edges.size(); c.winding = o.winding;
        if (o.geom == renderobject::GStrokedPath) { c.winding = COH_NONZERO; c.sprite_winding = 1 + COH_EVENODD; }  // render.ml:510, 1018
        if (o.convolved.kind) c.convolve = o.convolved.kind | (o.convolved.radius << 8);
    }
    objs.push_back(c);
  }
};

// The `view` of render_frame: scene, pages @ background (render.ml:1345-1365), resident on the device.
class view {
 public:
  view(const scene& s, const scene& background) {
    Flattened f;
    f.add_scene(s);
    f.add_background(background);
    ck(coh_scene_create(Context::get().raw(), f.objs.data(), (int32_t)f.objs.size(), f.n_background, f.edges.data(), (int32_t)(f.edges.size() / 4),
                        f.points.data(), (int32_t)(f.points.size() / 2), &h_));
  }
  view(const view&) = delete;
  view& operator=(const view&) = delete;
  ~view() { coh_scene_free(Context::get().raw(), h_); }
  coh_scene_t handle() const { return h_; }

 private:
  coh_scene_t h_ = 0;
};
// the canvas (wxgui.ml:254-262), now a device framebuffer
inline void set_canvas(int width, int height) { ck(coh_fb_configure(Context::get().raw(), width, height, 0, height)); }
// Render.render_frame lmo view update (render.mli:211-214), update = Sprite.box x y w h or any shape
inline void render_frame(const view& v, int x, int y, int w, int h) { ck(coh_render_frame(Context::get().raw(), v.handle(), x, y, w, h, COH_RENDER_RECORD_U)); }
inline void render_frame(const view& v, const Sprite::shape& update) { ck(coh_render_frame_shape(Context::get().raw(), v.handle(), update.handle(), COH_RENDER_RECORD_U)); }
// Render.render_simple_scene scene shape (render.mli:216-217)
inline void render_simple_scene(const scene& s, const Sprite::shape& shp) {
  view v(s, scene());
  render_frame(v, shp);
  ck(coh_sync(Context::get().raw()));
}
// the `u` render_scene returns (render.ml:1310-1335)
inline Sprite::shape uncovered() {
  coh_shape_t o = 0;
  ck(coh_render_uncovered(Context::get().raw(), &o));
  return Sprite::shape(o);
}
// Render.plaindirty / alldirty (render.ml:1376-1391)
inline Sprite::shape plaindirty(const Sprite::shape& so, const Sprite::shape& mo, const Sprite::shape& sn, const Sprite::shape& mn, const Sprite::shape& u) {
  coh_shape_t o = 0;
  ck(coh_dirty_region(Context::get().raw(), so.handle(), mo.handle(), sn.handle(), mn.handle(), u.handle(), 1, &o));
  return Sprite::shape(o);
}
inline Sprite::shape alldirty(const Sprite::shape& so, const Sprite::shape& sn, const Sprite::shape& u) {
  coh_shape_t o = 0;
  ck(coh_dirty_region(Context::get().raw(), so.handle(), 0, sn.handle(), 0, u.handle(), 0, &o));
  return Sprite::shape(o);
}
// Wxgui.plot_sprite's canvas bytes (wxgui.ml:417-424) and the RGBA8 framebuffer
inline std::vector<uint32_t> read_rgba(int x, int y, int w, int h) {
  std::vector<uint32_t> out((size_t)w * h);
  ck(coh_fb_read_rgba(Context::get().raw(), x, y, w, h, (uint8_t*)out.data()));
  return out;
}
inline std::vector<uint8_t> read_rgb888(int x, int y, int w, int h) {
  std::vector<uint8_t> out((size_t)w * h * 3);
  ck(coh_fb_read_rgb888(Context::get().raw(), x, y, w, h, out.data()));
  return out;
}
}  // namespace Render
// Wxgui.refresh_window window (xmin, ymin, xmax, ymax) (wxgui.ml:352-366): the marshalled "RefreshWindow" message of a
// dirty rectangle with the framebuffer's pixels, ready for the socket; empty where the reference sends nothing.
namespace Wxgui {
inline std::string refresh_window_message(int window, int xmin, int ymin, int xmax, int ymax) {
  int64_t n = 0;
  ck(coh_wire_refresh_window(Context::get().raw(), window, xmin, ymin, xmax, ymax, nullptr, 0, &n));
  std::string out((size_t)n, '\0');
  if (n > 0) ck(coh_wire_refresh_window(Context::get().raw(), window, xmin, ymin, xmax, ymax, (uint8_t*)&out[0], n, &n));
  return out;
}
}  // namespace Wxgui
}  // namespace coherence
