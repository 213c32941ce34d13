// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement
// of the reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product library never does.
// PARITY UNPINNED (no reference tests / golden vectors exist; OCaml cannot be built here).
// One exception: camlpy.hpp (the socket format) IS pinned — by vectors the reference's own pycaml.py produced here
// (tools/make_wire_golden.py -> tests/golden/wire_pycaml.json -> tests/test_wire.py).
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include "../include/coherence_b200.h"
#include "render.hpp"
#include "shapes.hpp"
#include "camlpy.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception& e) { g_err = e.what(); return 1; } return 0;

static std::vector<Edge> edges_from(const int32_t* e, int n) {
  std::vector<Edge> v((size_t)n);
  for (int i = 0; i < n; i++) v[i] = Edge{e[4 * i], e[4 * i + 2], e[4 * i + 1], e[4 * i + 3]};  // x0,y0,x1,y1 -> {x0,x1,y0,y1}
  return v;
}
static int32_t* dup_ints(const std::vector<int>& v) {
  int32_t* p = (int32_t*)std::malloc(sizeof(int32_t) * (v.size() + 1));
  if (!v.empty()) std::memcpy(p, v.data(), sizeof(int32_t) * v.size());
  return p;
}
static Fill fill_from(const coh_object& o) {
  Fill f;
  f.kind = (Fill::Kind)o.fill_kind;
  f.c0 = colour_of_rgba8(o.colour0); f.c1 = colour_of_rgba8(o.colour1);
  for (int i = 0; i < 6; i++) f.p[i] = o.fparam[i];
  f.ext_s = o.fill_flags & COH_FILL_EXT_S; f.ext_e = o.fill_flags & COH_FILL_EXT_E;
  return f;
}
// Rebuild the object tree from the flattened list.  An alias (dx,dy) is applied the way
// the reference's cache serves it: the ORIGINAL geometry's shape/sprite translated by whole
// pixels (cache.ml:380-385,400-405) — see Obj::dx in render.hpp.
static Scene build_scene(const coh_object* objs, int n, const int32_t* edges, const int32_t* points) {
  std::vector<Scene> stack(1);
  std::vector<Obj> open;
  std::vector<int> open_reading;                          // abi index of an open reading-scene group, or -1
  std::vector<int> open_conv;                             // per open group: its `convolve` word (Convolved (kernel, Group members))
  std::map<int, std::shared_ptr<Scene>> reading;          // reading-scene groups by the index of their GROUP_BEGIN
  std::vector<std::pair<size_t, int>> pending_reading;    // (position of the filter in the top-level list, first2)
  std::vector<std::pair<int, size_t>> awaiting_geom;      // COH_GEOM_NEXT filters: (stack depth, position in that list) waiting for the next object
  // the object just appended to the list at this depth is the geometry of the filter in front of it
  auto take_geometry = [&]() {
    if (awaiting_geom.empty() || awaiting_geom.back().first != (int)stack.size()) return;
    Scene& l = stack.back();
    const size_t fp = awaiting_geom.back().second;
    if (l.size() != fp + 2) return;
    Obj g = std::move(l.back()); l.pop_back();
    l[fp].children.push_back(std::move(g));
    awaiting_geom.pop_back();
  };
  for (int i = 0; i < n; i++) {
    const coh_object& c = objs[i];
    Obj o;
    o.id = c.id; o.pretrans = c.pretrans; o.dx = c.dx; o.dy = c.dy;
    for (int k = 0; k < 4; k++) o.bounds[k] = c.bounds[k];
    o.has_bounds = !(c.bounds[0] == 0 && c.bounds[1] == 0 && c.bounds[2] == 0 && c.bounds[3] == 0);
    switch (c.kind) {
      case COH_OBJ_PATH: {
        o.kind = Obj::Path; o.fill = fill_from(c); o.winding = (Winding)c.winding;
        o.sprite_winding = c.sprite_winding ? (Winding)(c.sprite_winding - 1) : o.winding;
        o.edges = edges_from(edges + 4 * (size_t)c.first, c.count);
        sort_edgelist_maxy_rev(o.edges);
        if (c.convolve) {  // Convolved (kernel, Basic (fill, Path p)): the path becomes the child
          Obj outer;
          outer.kind = Obj::Convolved; outer.id = c.id; outer.pretrans = c.pretrans; outer.dx = c.dx; outer.dy = c.dy;
          outer.fill = o.fill;
          int r = c.convolve >> 8;
          outer.kernel = (c.convolve & 255) == COH_CONV_UNIT ? mkunit(r) : mkgaussian(r);
          o.id = -1; o.pretrans = -1; o.dx = o.dy = 0; o.has_bounds = false;
          outer.children.push_back(std::move(o));
          stack.back().push_back(std::move(outer));
          break;
        }
        stack.back().push_back(std::move(o));
        break;
      }
      case COH_OBJ_FILTER: {
        o.kind = Obj::Filter; o.filter_kind = c.filter_kind; o.has_bounds = false;
        if (c.filter_kind == COH_FILTER_BLUR) o.kernel = (c.filter_kernel & 255) == COH_CONV_UNIT ? mkunit(c.filter_kernel >> 8) : mkgaussian(c.filter_kernel >> 8);
        Obj g;
        if (c.filter_kind != COH_FILTER_SMEAR && c.cpg_op == COH_GEOM_NEXT) {
          awaiting_geom.push_back({(int)stack.size(), stack.back().size()});   // the object that follows becomes children[0]
        } else if (c.filter_kind == COH_FILTER_SMEAR) {
          // geometry = Basic (white, Brushstroke (Brush.mkdummy brushstroke)) (filters.ml:205-207)
          o.stroke.opacity = c.brush_opacity; o.stroke.radius = c.brush_radius;
          for (int k = 0; k < c.count; k++) o.stroke.points.push_back({points[2 * ((size_t)c.first + k)], points[2 * ((size_t)c.first + k) + 1]});
          for (int k = 0; k < c.count2; k++) o.smear_pts.push_back({points[2 * ((size_t)c.first2 + k)], points[2 * ((size_t)c.first2 + k) + 1]});
          g.kind = Obj::Brush; g.fill = Fill::plain(mkcol(255, 255, 255));
          g.stroke = o.stroke; g.stroke.dummy = true; g.stroke.rx = (o.stroke.bw() - 1) / 2; g.stroke.opacity = 1.;
        } else {
          g.kind = Obj::Path; g.fill = fill_from(c); g.winding = g.sprite_winding = (Winding)c.winding;
          g.edges = edges_from(edges + 4 * (size_t)c.first, c.count);
          sort_edgelist_maxy_rev(g.edges);
        }
        if (!(c.filter_kind != COH_FILTER_SMEAR && c.cpg_op == COH_GEOM_NEXT)) o.children.push_back(std::move(g));
        // (a filter inside a Group sees the rest of that group's list as its objects below, render.ml:988-1001)
        if (c.filter_kind == COH_FILTER_SCENE) {
          if (stack.size() != 1) throw std::runtime_error("scene: a filter with a caller-built reading scene must be top-level");
          pending_reading.push_back({stack.back().size(), c.first2});
        }
        stack.back().push_back(std::move(o));
        break;
      }
      case COH_OBJ_CPG: {
        o.kind = Obj::CPG; o.fill = fill_from(c); o.cpg_op = c.cpg_op;
        Obj a, b;
        a.kind = b.kind = Obj::Path;
        a.winding = a.sprite_winding = (Winding)c.winding; b.winding = b.sprite_winding = (Winding)c.winding2;
        a.edges = edges_from(edges + 4 * (size_t)c.first, c.count); b.edges = edges_from(edges + 4 * (size_t)c.first2, c.count2);
        sort_edgelist_maxy_rev(a.edges); sort_edgelist_maxy_rev(b.edges);
        o.has_bounds = false;
        o.children.push_back(std::move(a)); o.children.push_back(std::move(b));
        stack.back().push_back(std::move(o));
        break;
      }
      case COH_OBJ_PRIMITIVE: {
        o.kind = Obj::Primitive; o.prim_colour = colour_of_rgba8(c.colour0); o.prim_null = c.prim_null;
        for (int k = 0; k < 4; k++) o.prim[k] = c.prim[k];
        stack.back().push_back(std::move(o));
        break;
      }
      case COH_OBJ_BRUSH: {
        o.kind = Obj::Brush; o.fill = fill_from(c);
        o.stroke.opacity = c.brush_opacity; o.stroke.radius = c.brush_radius;
        if (c.winding == COH_BRUSH_DUMMY) { o.stroke.dummy = true; o.stroke.rx = (int)c.brush_radius; }
        for (int k = 0; k < c.count; k++)
          o.stroke.points.push_back({points[2 * ((size_t)c.first + k)], points[2 * ((size_t)c.first + k) + 1]});
        stack.back().push_back(std::move(o));
        break;
      }
      case COH_OBJ_GROUP_BEGIN: {
        o.kind = Obj::Group;
        if (c.filter_kind == COH_FILTER_READING_SCENE && !open.empty()) throw std::runtime_error("scene: reading-scene groups must be top-level");
        open_reading.push_back(c.filter_kind == COH_FILTER_READING_SCENE ? i : -1);
        open.push_back(std::move(o)); stack.emplace_back(); open_conv.push_back(c.convolve);
        break;
      }
      case COH_OBJ_GROUP_END: {
        if (open.empty()) throw std::runtime_error("scene: GROUP_END without GROUP_BEGIN");
        Obj g = std::move(open.back()); open.pop_back();
        g.children = std::move(stack.back()); stack.pop_back();
        const int rd = open_reading.back(); open_reading.pop_back();
        const int conv = open_conv.back(); open_conv.pop_back();
        if (rd >= 0) { reading[rd] = std::make_shared<Scene>(std::move(g.children)); break; }
        if (g.children.empty()) throw std::runtime_error("Empty groups aren't allowed");  // render.ml:317
        if (conv) {  // Convolved (kernel, Group members) (render.ml:63, 1023-1052): the group becomes the child
          Obj outer;
          outer.kind = Obj::Convolved; outer.id = g.id; outer.pretrans = g.pretrans; outer.dx = g.dx; outer.dy = g.dy;
          const int r = conv >> 8;
          outer.kernel = (conv & 255) == COH_CONV_UNIT ? mkunit(r) : mkgaussian(r);
          g.id = -1; g.pretrans = -1; g.dx = g.dy = 0; g.has_bounds = false;
          outer.children.push_back(std::move(g));
          stack.back().push_back(std::move(outer));
          break;
        }
        stack.back().push_back(std::move(g));
        break;
      }
      default: throw std::runtime_error("scene: unknown object kind");
    }
    take_geometry();
  }
  if (!awaiting_geom.empty()) throw std::runtime_error("scene: a filter with COH_GEOM_NEXT is followed by its geometry object");
  if (!open.empty()) throw std::runtime_error("scene: unterminated group");
  for (auto& pr : pending_reading) {
    auto it = reading.find(pr.second);
    if (it == reading.end()) throw std::runtime_error("scene: filter without its reading-scene group");
    stack[0].at(pr.first).reading_scene = it->second;
  }
  return stack[0];
}
static void sprite_to_dense(const Sprite& s, int ux, int uy, int uw, int uh, uint32_t* out) {
  for (auto& r : s.rows) {
    if (r.y < uy || r.y >= uy + uh) continue;
    int off = 0;
    for (auto& sp : r.spans) {
      for (int k = 0; k < sp.len; k++) {
        int x = sp.x + k;
        if (x >= ux && x < ux + uw) out[(size_t)(r.y - uy) * uw + (x - ux)] = rgba8_of_colour(r.px[off + k]);
      }
      off += sp.len;
    }
  }
}

extern "C" {
// Brush.points_of_brushstroke_smear + the integer points of find_smear_directions (brush.ml:239-283); segs = records of
// 9 doubles (kind 0 straight / 1 bezier, then up to 4 points), all segments of the path in order
int64_t orc_smear_points(const double* segs, int32_t n_segs, int32_t* out, int64_t cap) {
  Path path(1);
  for (int i = 0; i < n_segs; i++) {
    const double* q = segs + 9 * i;
    Segment sg; sg.bezier = q[0] != 0.;
    for (int k = 0; k < 4; k++) sg.p[k] = Pt(q[1 + 2 * k], q[2 + 2 * k]);
    path[0].push_back(sg);
  }
  auto ip = smear_int_points(points_of_brushstroke_smear(path));
  for (size_t i = 0; i < ip.size() && (int64_t)i < cap; i++) { out[2 * i] = ip[i].first; out[2 * i + 1] = ip[i].second; }
  return (int64_t)ip.size();
}
const char* orc_last_error() { return g_err.c_str(); }
void orc_free(void* p) { std::free(p); }

int32_t orc_colour_of_rgba8(uint32_t w) { return colour_of_rgba8(w); }
uint32_t orc_rgba8_of_colour(int32_t c) { return rgba8_of_colour(c); }
int32_t orc_colour_of_rgba(int r, int g, int b, int a) { return colour_of_rgba(r, g, b, a); }
int orc_div255(int i) { return div255(i); }
// op: 0 over, 1 alpha_over, 2 pd_plus, 3 dissolve(a, delta=b), 4 dissolve_between(a,b,alpha=c), 5 monochrome(a)
int orc_colour_op(int op, uint32_t a, uint32_t b, int c, uint32_t* out) {
  ORC_TRY
  colour ca = colour_of_rgba8(a), cb = colour_of_rgba8(b), r;
  switch (op) {
    case 0: r = over(ca, cb); break;
    case 1: r = alpha_over(ca, cb); break;
    case 2: r = pd_plus(ca, cb); break;
    case 3: r = dissolve(ca, (int)b); break;
    case 4: r = dissolve_between(ca, cb, c); break;
    case 5: r = monochrome(ca); break;
    default: throw std::runtime_error("orc_colour_op: bad op");
  }
  *out = rgba8_of_colour(r);
  ORC_CATCH
}
int orc_sub_of_float(double f) { return sub_of_float(f); }
int orc_pix_of_sub(int n) { return pix_of_sub(n); }

// AA tables: maintable (32*32 ints, [x][y]) and volume.
int orc_aa_tables(int32_t* maintable, int32_t* volume) {
  const AATables& T = aa_tables();
  for (int x = 0; x < 32; x++) for (int y = 0; y < 32; y++) maintable[x * 32 + y] = T.maintable[x][y];
  *volume = T.volume;
  return 0;
}

// Polygon.shapeminshape_of_unsorted_edgelist: returns two malloc'd flat shapes.
int orc_shapeminshape(const int32_t* edges, int n, int winding, int32_t** shp, int64_t* nshp, int32_t** minshp, int64_t* nmin) {
  ORC_TRY
  Shape s, m;
  shapeminshape_of_unsorted_edgelist(edges_from(edges, n), (Winding)winding, s, m);
  if (!shapecheck(s) || !shapecheck(m)) throw std::runtime_error("shapeminshape: malformed output");
  auto a = shape_to_flat(s), b = shape_to_flat(m);
  *shp = dup_ints(a); *nshp = (int64_t)a.size(); *minshp = dup_ints(b); *nmin = (int64_t)b.size();
  ORC_CATCH
}
// The x16 scaled shape of polygon.ml:673-692 (debug / cross-check artefact).
int orc_scaled_shape(const int32_t* edges, int n, int winding, int32_t** shp, int64_t* nshp) {
  ORC_TRY
  std::vector<Edge> es = edges_from(edges, n);
  sort_edgelist_maxy_rev(es);
  auto a = shape_to_flat(mk_scaled_shape((Winding)winding, es));
  *shp = dup_ints(a); *nshp = (int64_t)a.size();
  ORC_CATCH
}
// AA opacity for every pixel of `shape` (flat), span order.
int orc_polygon_opacity(const int32_t* edges, int n, int winding, const int32_t* shape, int64_t nshape, uint8_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  std::vector<Edge> es = edges_from(edges, n);
  sort_edgelist_maxy_rev(es);
  Shape scaled = mk_scaled_shape((Winding)winding, es);
  Shape shp = shape_from_flat(shape, (int)nshape);
  int64_t k = 0;
  for (auto& r : shp.rows) for (auto& sp : r.spans) for (int i = 0; i < sp.len; i++) {
    if (k >= cap) throw std::runtime_error("orc_polygon_opacity: buffer too small");
    out[k++] = (uint8_t)pixel_opacity(scaled, sp.x + i, r.y);
  }
  *nout = k;
  ORC_CATCH
}
// Polygon.polygon_sprite_edgelist: RGBA8 per pixel of shape, span order.
int orc_polygon_sprite(const coh_object* fill, const int32_t* edges, int n, int winding, const int32_t* shape, int64_t nshape, uint32_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  std::vector<Edge> es = edges_from(edges, n);
  sort_edgelist_maxy_rev(es);
  Sprite s = Renderer::polygon_sprite_edgelist(fill_from(*fill), shape_from_flat(shape, (int)nshape), es, (Winding)winding);
  int64_t k = 0;
  for (auto& r : s.rows) for (colour c : r.px) {
    if (k >= cap) throw std::runtime_error("orc_polygon_sprite: buffer too small");
    out[k++] = rgba8_of_colour(c);
  }
  *nout = k;
  ORC_CATCH
}
// op: 0 union 1 difference 2 intersection
int orc_shape_op(int op, const int32_t* a, int64_t na, const int32_t* b, int64_t nb, int32_t** out, int64_t* nout) {
  ORC_TRY
  Shape A = shape_from_flat(a, (int)na), B = shape_from_flat(b, (int)nb), R;
  if (!shapecheck(A) || !shapecheck(B)) throw std::runtime_error("shape op: malformed input");
  R = op == 0 ? shape_union(A, B) : op == 1 ? shape_difference(A, B) : shape_intersection(A, B);
  if (!shapecheck(R)) throw std::runtime_error("shape op: malformed output");
  auto f = shape_to_flat(R); *out = dup_ints(f); *nout = (int64_t)f.size();
  ORC_CATCH
}
// op: 0 bloat m n, 1 erode m n, 2 translate dx dy
int orc_shape_unary(int op, const int32_t* a, int64_t na, int m, int n, int32_t** out, int64_t* nout) {
  ORC_TRY
  Shape A = shape_from_flat(a, (int)na), R;
  R = op == 0 ? bloat(m, n, A) : op == 1 ? erode(m, n, A) : translate_shape(m, n, A);
  if (!shapecheck(R)) throw std::runtime_error("shape unary: malformed output");
  auto f = shape_to_flat(R); *out = dup_ints(f); *nout = (int64_t)f.size();
  ORC_CATCH
}

// Render.render_frame over update = Sprite.box ux uy uw uh.  `out` is a dense uw*uh RGBA8
// image (0 where the result sprite has no pixel).  If u_out != NULL it receives the flat
// shape of the covered-so-far complement `u` after the scene pass.
// flags bit0: disable the bbox trivial reject.  usecache: Cache.usecache.
int orc_render_frame(const coh_object* objs, int n_scene, int n_background, const int32_t* edges, const int32_t* points,
                     int ux, int uy, int uw, int uh, int flags, int usecache, uint32_t* out, int32_t** u_out, int64_t* nu_out) {
  ORC_TRY
  Scene scene = build_scene(objs, n_scene, edges, points);
  Scene bg = build_scene(objs + n_scene, n_background, edges, points);
  Renderer R; R.bbox_reject = !(flags & 1); R.cache.usecache = usecache != 0;
  Shape update = shape_box(ux, uy, uw, uh);
  std::memset(out, 0, sizeof(uint32_t) * (size_t)uw * uh);
  Shape u1 = update; Sprite a1; R.render_scene(u1, a1, scene, false);
  Sprite res = a1;
  if (n_background > 0) {
    Shape u2 = update; Sprite a2; R.render_scene(u2, a2, bg, true);
    res = caf(over, opaque, a1, a2).first;
  }
  sprite_to_dense(res, ux, uy, uw, uh, out);
  if (u_out) { auto f = shape_to_flat(u1); *u_out = dup_ints(f); *nu_out = (int64_t)f.size(); }
  ORC_CATCH
}

// Convolve.convolve_sprite (kind 1 = mkunit r, 2 = mkgaussian r): sprite = flat shape + RGBA8 per pixel.
int orc_convolve_sprite(int kind, int r, const int32_t* shape, int64_t nshape, const uint32_t* rgba, int32_t** out_shape, int64_t* n_out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  ORC_TRY
  Shape shp = shape_from_flat(shape, (int)nshape);
  Sprite spr;
  int64_t k = 0;
  for (auto& row : shp.rows) {
    SpriteRow sr; sr.y = row.y; sr.spans = row.spans;
    for (auto& sp : row.spans) for (int i = 0; i < sp.len; i++) sr.px.push_back(colour_of_rgba8(rgba[k++]));
    spr.rows.push_back(std::move(sr));
  }
  Sprite res = convolve_sprite(kind == 1 ? mkunit(r) : mkgaussian(r), spr);
  auto f = shape_to_flat(shape_of_sprite(res));
  *out_shape = dup_ints(f); *n_out_shape = (int64_t)f.size();
  int64_t n = 0;
  for (auto& row : res.rows) for (colour c : row.px) { if (n >= cap) throw std::runtime_error("orc_convolve_sprite: buffer too small"); rgba_out[n++] = rgba8_of_colour(c); }
  *n_out = n;
  ORC_CATCH
}

// Sprite operations on (flat shape, RGBA8 per pixel in span order) sprites.
static Sprite sprite_from(const int32_t* shape, int64_t nshape, const uint32_t* rgba) {
  Shape shp = shape_from_flat(shape, (int)nshape);
  Sprite spr; int64_t k = 0;
  for (auto& row : shp.rows) {
    SpriteRow sr; sr.y = row.y; sr.spans = row.spans;
    for (auto& sp : row.spans) for (int i = 0; i < sp.len; i++) sr.px.push_back(colour_of_rgba8(rgba[k++]));
    spr.rows.push_back(std::move(sr));
  }
  return spr;
}
static int64_t sprite_out(const Sprite& s, uint32_t* out, int64_t cap) {
  int64_t n = 0;
  for (auto& row : s.rows) for (colour c : row.px) { if (n >= cap) throw std::runtime_error("sprite buffer too small"); out[n++] = rgba8_of_colour(c); }
  return n;
}
// Brush strokes outside a scene (brush.mli:20-27).  brush: a BRUSH object record; points: rounded stamp points.
static BrushStroke stroke_from(const coh_object* b, const int32_t* points, int n) {
  BrushStroke st; st.opacity = b->brush_opacity; st.radius = b->brush_radius;
  if (b->winding == COH_BRUSH_DUMMY) { st.dummy = true; st.rx = (int)b->brush_radius; }
  for (int k = 0; k < n; k++) st.points.push_back({points[2 * k], points[2 * k + 1]});
  return st;
}
int orc_brush_shape(const coh_object* brush, const int32_t* points, int n, int32_t** out, int64_t* nout) {
  ORC_TRY
  auto f = shape_to_flat(shape_of_brushstroke(stroke_from(brush, points, n)));
  *out = dup_ints(f); *nout = (int64_t)f.size();
  ORC_CATCH
}
int orc_brush_sprite(const coh_object* brush, const int32_t* points, int n, const int32_t* shape, int64_t nshape, uint32_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  *nout = sprite_out(sprite_of_brushstroke(stroke_from(brush, points, n), fill_from(*brush), shape_from_flat(shape, (int)nshape)), out, cap);
  ORC_CATCH
}
int orc_brush_smear(const int32_t* shape, int64_t nshape, const uint32_t* rgba, const coh_object* brush, const int32_t* points, int n,
                    const int32_t* smear_points, int n_smear, int32_t** out_shape, int64_t* n_out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out) {
  ORC_TRY
  std::vector<std::pair<int, int>> sp;
  for (int k = 0; k < n_smear; k++) sp.push_back({smear_points[2 * k], smear_points[2 * k + 1]});
  Sprite in = nshape ? sprite_from(shape, nshape, rgba) : Sprite();
  Sprite res = smear(in, stroke_from(brush, points, n), sp);
  auto f = shape_to_flat(shape_of_sprite(res));
  *out_shape = dup_ints(f); *n_out_shape = (int64_t)f.size();
  *n_out = sprite_out(res, rgba_out, cap);
  ORC_CATCH
}
// Sprite.portion spr shp (sprite.ml:642-721)
int orc_sprite_portion(const int32_t* shape, int64_t nshape, const uint32_t* rgba, const int32_t* sub, int64_t nsub, uint32_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  *nout = sprite_out(portion(sprite_from(shape, nshape, rgba), shape_from_flat(sub, (int)nsub)), out, cap);
  ORC_CATCH
}
// Sprite.fillshape shp fill (sprite.ml:158-175)
int orc_sprite_fillshape(const coh_object* fill, const int32_t* shape, int64_t nshape, uint32_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  *nout = sprite_out(fillshape(shape_from_flat(shape, (int)nshape), fill_from(*fill)), out, cap);
  ORC_CATCH
}
// Sprite.sprite_map f (sprite.ml:358-374) with f = 0 Colour.monochrome | 1 dissolve ~delta:arg | 2 / 3 / 4 red / green / blue_channel
int orc_sprite_map(int op, int arg, const uint32_t* rgba, int64_t n, uint32_t* out) {
  ORC_TRY
  for (int64_t i = 0; i < n; i++) {
    colour c = colour_of_rgba8(rgba[i]);
    colour r = op == 0 ? monochrome(c) : op == 1 ? dissolve(c, arg) : op == 2 ? red_channel(c) : op == 3 ? green_channel(c) : blue_channel(c);
    out[i] = rgba8_of_colour(r);
  }
  ORC_CATCH
}
// Sprite.map_coords (fun x y c -> dissolve (fill x y) (alpha c)) (render.ml:976-981, the last step of sprite_of_cpg)
int orc_sprite_map_coords_fill(const coh_object* fill, const int32_t* shape, int64_t nshape, const uint32_t* rgba, uint32_t* out, int64_t cap, int64_t* nout) {
  ORC_TRY
  Sprite s = sprite_from(shape, nshape, rgba);
  Fill f = fill_from(*fill);
  for (auto& r : s.rows) {
    int off = 0;
    for (auto& sp : r.spans) { for (int k = 0; k < sp.len; k++) r.px[off + k] = dissolve(f.fillsingle(sp.x + k, r.y), alpha_of_colour(r.px[off + k])); off += sp.len; }
  }
  *nout = sprite_out(s, out, cap);
  ORC_CATCH
}

// Persistent renderer for the cached / animated configurations (C4): keeps Cache between frames.
void* orc_renderer_new(int usecache) { Renderer* r = new Renderer(); r->cache.usecache = usecache != 0; return r; }
void orc_renderer_free(void* r) { delete (Renderer*)r; }
int orc_renderer_addtranslation(void* r, int64_t id, int64_t target, int dx, int dy) {
  ((Renderer*)r)->cache.addtranslation(id, target, dx, dy); return 0;
}
int orc_renderer_frame(void* rp, const coh_object* objs, int n_scene, int n_background, const int32_t* edges, const int32_t* points,
                       const int32_t* update_flat, int64_t n_update, int ox, int oy, int ow, int oh, uint32_t* out) {
  ORC_TRY
  Renderer& R = *(Renderer*)rp;
  Scene scene = build_scene(objs, n_scene, edges, points);
  Scene bg = build_scene(objs + n_scene, n_background, edges, points);
  Shape update = shape_from_flat(update_flat, (int)n_update);
  Sprite res = R.render_frame(scene, bg, update);
  sprite_to_dense(res, ox, oy, ow, oh, out);
  ORC_CATCH
}

// Host-side geometry helpers restated from polygon.ml (used to cross-check the product's
// own host-side flattening): bezier flattening and points_on_path.
int orc_flatten_bezier(const double* p8, double eps, double** out, int64_t* nseg) {
  ORC_TRY
  std::vector<std::pair<Pt, Pt>> e;
  bezier_subdivide(eps, Pt(p8[0], p8[1]), Pt(p8[2], p8[3]), Pt(p8[4], p8[5]), Pt(p8[6], p8[7]), e);
  double* o = (double*)std::malloc(sizeof(double) * 4 * (e.size() + 1));
  for (size_t i = 0; i < e.size(); i++) { o[4 * i] = e[i].first.first; o[4 * i + 1] = e[i].first.second; o[4 * i + 2] = e[i].second.first; o[4 * i + 3] = e[i].second.second; }
  *out = o; *nseg = (int64_t)e.size();
  ORC_CATCH
}
// segs: nseg records of 9 doubles: kind (0 straight, 1 bezier), then 4 points (straight uses 2).
// One subpath.  Returns rounded integer points (brush.ml:172).
int orc_points_on_path(const double* segs, int nseg, double sep, int32_t** out, int64_t* npts) {
  ORC_TRY
  Subpath sub;
  for (int i = 0; i < nseg; i++) {
    Segment s; s.bezier = segs[9 * i] != 0.;
    for (int k = 0; k < 4; k++) s.p[k] = Pt(segs[9 * i + 1 + 2 * k], segs[9 * i + 2 + 2 * k]);
    sub.push_back(s);
  }
  auto pts = round_points(points_on_path(sep, Path{sub}));
  std::vector<int> f;
  for (auto& p : pts) { f.push_back(p.first); f.push_back(p.second); }
  *out = dup_ints(f); *npts = (int64_t)pts.size();
  ORC_CATCH
}
// Shapes.strokepath_polygon / Shapes.strokepath (shapes.ml:203-530).  spec5: startcap, join, endcap (oracle enums: caps
// Butt 0 / Round 1 / Projecting 2, joins Round 0 / Mitred 1 / Bevel 2), mitrelimit, linewidth.  subpath_n: segments per
// subpath.  Returns the outline (9-double records, segments per outline subpath), its winding rule and the sorted edges.
int orc_strokepath(const double* spec5, const double* segs, const int32_t* subpath_n, int n_subpaths, double** segs_out, int64_t* n_segs_out,
                   int32_t** subpath_n_out, int64_t* n_subpaths_out, int* winding_out, int32_t** edges_out, int64_t* n_edges_out) {
  ORC_TRY
  StrokeSpec spec{(Cap)(int)spec5[0], (Join)(int)spec5[1], (Cap)(int)spec5[2], spec5[3], spec5[4]};
  Path path;
  int at = 0;
  for (int k = 0; k < n_subpaths; k++) {
    Subpath sub;
    for (int i = 0; i < subpath_n[k]; i++, at++) {
      Segment s; s.bezier = segs[9 * at] != 0.;
      for (int q = 0; q < 4; q++) s.p[q] = Pt(segs[9 * at + 1 + 2 * q], segs[9 * at + 2 + 2 * q]);
      sub.push_back(s);
    }
    path.push_back(sub);
  }
  Path outline;
  *winding_out = (int)strokepath_polygon(spec, path, outline);
  std::vector<double> f; std::vector<int> cnt;
  for (const Subpath& sp : outline) {
    cnt.push_back((int)sp.size());
    for (const Segment& s : sp) {
      f.push_back(s.bezier ? 1. : 0.);
      for (int q = 0; q < 4; q++) { f.push_back(s.bezier || q < 2 ? s.p[q].first : 0.); f.push_back(s.bezier || q < 2 ? s.p[q].second : 0.); }
    }
  }
  double* o = (double*)std::malloc(sizeof(double) * (f.size() + 1));
  std::copy(f.begin(), f.end(), o);
  *segs_out = o; *n_segs_out = (int64_t)(f.size() / 9);
  *subpath_n_out = dup_ints(cnt); *n_subpaths_out = (int64_t)cnt.size();
  std::vector<Edge> es = strokepath(spec, path);
  std::vector<int> e;
  for (const Edge& d : es) { e.push_back(d.x0); e.push_back(d.y0); e.push_back(d.x1); e.push_back(d.y1); }
  *edges_out = dup_ints(e); *n_edges_out = (int64_t)es.size();
  ORC_CATCH
}
int orc_bounds_stroke(const double* spec5, const double* segs, const int32_t* subpath_n, int n_subpaths, int32_t* bounds4) {
  ORC_TRY
  StrokeSpec spec{(Cap)(int)spec5[0], (Join)(int)spec5[1], (Cap)(int)spec5[2], spec5[3], spec5[4]};
  Path path;
  int at = 0;
  for (int k = 0; k < n_subpaths; k++) {
    Subpath sub;
    for (int i = 0; i < subpath_n[k]; i++, at++) {
      Segment s; s.bezier = segs[9 * at] != 0.;
      for (int q = 0; q < 4; q++) s.p[q] = Pt(segs[9 * at + 1 + 2 * q], segs[9 * at + 2 + 2 * q]);
      sub.push_back(s);
    }
    path.push_back(sub);
  }
  int b[4];
  bounds_stroke(path, spec, b);
  for (int k = 0; k < 4; k++) bounds4[k] = b[k];
  ORC_CATCH
}
int orc_brush_stamp(double radius, double opacity, uint8_t* alpha_out, int cap, int* size_out) {
  ORC_TRY
  int size; auto b = drawround(radius, opacity, mkcol(255, 255, 255), size);
  if (size * size > cap) throw std::runtime_error("orc_brush_stamp: buffer too small");
  for (int i = 0; i < size * size; i++) alpha_out[i] = (uint8_t)alpha_of_colour(b[i]);
  *size_out = size;
  ORC_CATCH
}
// ---- camlpy.ml (oracle/camlpy.hpp): a marshallable crosses as the pre-order token list of include/coherence_b200.h ----
static bool tree_of_tokens(const int32_t* kinds, const int64_t* values, const int64_t* offsets, int n, const uint8_t* strings, int& at, camlpy::Marshallable& m) {
  if (at >= n) return false;
  const int i = at++;
  switch (kinds[i]) {
    case 1: m.kind = camlpy::Marshallable::Unit; return true;
    case 2: m.kind = camlpy::Marshallable::Int; m.i = values[i]; return true;
    case 4: m.kind = camlpy::Marshallable::Bool; m.i = values[i] != 0; return true;
    case 3: m.kind = camlpy::Marshallable::String; if (values[i] < 0) return false; m.s.assign((const char*)strings + offsets[i], (size_t)values[i]); return true;
    case 0:
      m.kind = camlpy::Marshallable::Tuple;
      if (values[i] < 0) return false;
      for (int64_t k = 0; k < values[i]; k++) { camlpy::Marshallable c; if (!tree_of_tokens(kinds, values, offsets, n, strings, at, c)) return false; m.members.push_back(c); }
      return true;
    default: return false;
  }
}
static void tokens_of_tree(const camlpy::Marshallable& m, int64_t& pos, std::vector<int32_t>& kinds, std::vector<int64_t>& values, std::vector<int64_t>& offsets) {
  kinds.push_back((int32_t)m.kind);
  switch (m.kind) {
    case camlpy::Marshallable::Unit: values.push_back(0); offsets.push_back(0); pos += 1; break;
    case camlpy::Marshallable::Int: values.push_back(m.i); offsets.push_back(0); pos += 5; break;
    case camlpy::Marshallable::Bool: values.push_back(m.i); offsets.push_back(0); pos += 2; break;
    case camlpy::Marshallable::String: values.push_back((int64_t)m.s.size()); offsets.push_back(pos + 5); pos += 5 + (int64_t)m.s.size(); break;
    default:
      values.push_back((int64_t)m.members.size()); offsets.push_back(0); pos += 5;
      for (const camlpy::Marshallable& c : m.members) tokens_of_tree(c, pos, kinds, values, offsets);
  }
}
// Camlpy.marshall: *size_out = -1 when the tokens are not exactly one value
int orc_wire_marshal(const int32_t* kinds, const int64_t* values, const int64_t* offsets, int n, const uint8_t* strings, uint8_t* out, int64_t cap, int64_t* size_out) {
  ORC_TRY
  camlpy::Marshallable m; int at = 0;
  if (n <= 0 || !tree_of_tokens(kinds, values, offsets, n, strings, at, m) || at != n) { *size_out = -1; return 0; }
  const std::string s = camlpy::marshall(m);
  *size_out = (int64_t)s.size();
  if (out && cap >= (int64_t)s.size()) std::memcpy(out, s.data(), s.size());
  ORC_CATCH
}
// Camlpy.unmarshall: *status = 0 None, 1 Some (taken, value), -1 Invalid_data; tokens up to cap, *n_tokens whatever cap is
int orc_wire_unmarshal(const uint8_t* buf, int64_t n, int32_t* kinds, int64_t* values, int64_t* offsets, int cap, int* n_tokens, int64_t* taken, int* status) {
  ORC_TRY
  *n_tokens = 0; *taken = 0; *status = 0;
  camlpy::Marshallable m; long long tk = 0;
  bool some;
  try { some = camlpy::unmarshall(std::string((const char*)buf, (size_t)n), tk, m); }
  catch (const camlpy::Invalid_data&) { *status = -1; return 0; }
  if (!some) return 0;
  std::vector<int32_t> k; std::vector<int64_t> v, o; int64_t pos = 4;
  tokens_of_tree(m, pos, k, v, o);
  *n_tokens = (int)k.size(); *taken = tk; *status = 1;
  for (int i = 0; i < (int)k.size() && i < cap; i++) { kinds[i] = k[i]; values[i] = v[i]; offsets[i] = o[i]; }
  ORC_CATCH
}
}  // extern "C"
