// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
// CPU restatement of the reference's scene renderer: per-object shapes, the
// min/max-shape split, the partial-sprite cache and the front-to-back loop with
// hidden-surface set subtraction.  Follows /root/reference/render.ml:19-75 (scene
// types), 469-594 (shape_of_basicshape), 984-1078 (sprite_of_basicshape), 1134-1242
// (spriteof), 1268-1335 (renderobj, render_scene), 1345-1370 (render_frame,
// render_simple_scene), 1376-1400 (dirty regions) and /root/reference/cache.ml:57-436.
// Transforms are applied before this boundary: geometry arrives as
// integer sub-bin edges, so no camlpdf arithmetic is involved.
// PARITY UNPINNED (no reference tests / golden vectors exist).
#pragma once
#include <map>
#include <memory>
#include "brush.hpp"
#include "convolve.hpp"
#include "polygon.hpp"
#include "sprite.hpp"
namespace oracle {

struct Obj;
typedef std::vector<Obj> Scene;

struct Obj {
  enum Kind { Path = 0, Primitive = 1, Group = 2, Brush = 3, Convolved = 4, CPG = 5, Filter = 6 } kind = Path;
  long id = -1;              // < 0: a fresh id per render (Id.new_ids ()), never cached
  int pretrans = -1;         // -1: Over; else PreTrans(v, Over) with delta = toint (v *. 255.)
  // Integer-pixel alias (Render.translate_renderobject -> Cache.addtranslation, render.ml:259-271,
  // cache.ml:423-436): shape and sprite are those of the geometry at its ORIGINAL position,
  // translated (cache.ml:380-385, 400-405).  This is not the same as rasterising moved edges where
  // coordinates go negative (pix_of_sub and toint truncate toward zero), so the alias is modelled
  // the way the cache serves it.
  int dx = 0, dy = 0;
  int bounds[4] = {0, 0, 0, 0};  // bounds_of_basicshape: xmin, xmax, ymin, ymax (render.ml:377-437)
  bool has_bounds = false;
  // Path
  Fill fill;
  std::vector<Edge> edges;   // sorted by sort_edgelist_maxy_rev
  Winding winding = NonZero;
  Winding sprite_winding = NonZero;  // render.ml:1018: stroked paths take their sprite with EvenOdd
  // Primitive (render.ml:556-586): inclusive integer box; prim_null for zero-length lines
  colour prim_colour = 0;
  int prim[4] = {0, 0, 0, 0};  // x0, y0, x1, y1 inclusive
  bool prim_null = false;
  // Group
  Scene children;
  // Brush
  BrushStroke stroke;
  // Convolved (kernel, child geometry = children[0])
  Kernel kernel;
  // CPG (op, a = children[0], b = children[1]), render.ml:17-18
  int cpg_op = 0;  // 0 Union, 1 Intersection, 2 Subtraction, 3 ExclusiveOr
  // Filter {geometry = children[0]; reading_scene; filter} (render.ml:37-48; filters.ml).  The
  // reference's closures are restated per filter kind: 1 hole, 2 monochrome, 3 blur (kernel),
  // 4 identity filter over a caller-built reading scene (affine / rgb / wireframe / swapdepth / minus).
  // 5 minus; 6 smear: `stroke` is the (transformed) brush stroke, the geometry children[0] its dummy (filters.ml:201-217),
  // smear_pts the deduplicated integer points of Brush.find_smear_directions (brush.ml:266-283)
  int filter_kind = 0;
  std::shared_ptr<Scene> reading_scene;
  std::vector<std::pair<int, int>> smear_pts;
};

// ---- cache.ml restated: shapes and partial sprites keyed by id, integer-translation
// aliases.  Eviction order in the reference follows Hashtbl.iter (unspecified); the
// cache never changes a rendered value, only the work done, so a simple byte budget
// with "drop sprites first, then shapes" stands in for drophalf (cache.ml:242-271).
struct Cache {
  struct Entry {
    bool alias = false; int dx = 0, dy = 0; long target = 0;
    bool has_shape = false; Shape shape, minshape;
    bool has_sprite = false; Sprite sprite; Shape pshape;
  };
  bool usecache = false;
  std::map<long, Entry> tab;
  long shphit = 0, shpmis = 0, sprhit = 0, sprmis = 0;
  void clear() { tab.clear(); }
  bool getshape(long id, Shape& s, Shape& m) {  // cache.ml:370-387
    if (!usecache || id < 0) return false;
    auto it = tab.find(id);
    if (it == tab.end()) { shpmis++; return false; }
    Entry& e = it->second;
    if (e.alias) {
      Shape s0, m0;
      if (!getshape(e.target, s0, m0)) return false;
      s = translate_shape(e.dx, e.dy, s0); m = translate_shape(e.dx, e.dy, m0);
      return true;
    }
    if (!e.has_shape) { shpmis++; return false; }
    shphit++; s = e.shape; m = e.minshape; return true;
  }
  void addshape(long id, const Shape& s, const Shape& m) {  // cache.ml:280-324
    if (!usecache || id < 0) return;
    auto it = tab.find(id);
    if (it != tab.end() && it->second.alias) { id = it->second.target; it = tab.find(id); }
    if (it != tab.end() && it->second.has_shape) return;
    Entry& e = tab[id];
    e.has_shape = true; e.shape = s; e.minshape = m;
  }
  bool getsprite(long id, Sprite& spr, Shape& pshape) {  // cache.ml:390-407
    if (!usecache || id < 0) return false;
    auto it = tab.find(id);
    if (it == tab.end()) { sprmis++; return false; }
    Entry& e = it->second;
    if (e.alias) {
      Sprite s0; Shape p0;
      if (!getsprite(e.target, s0, p0)) return false;
      spr = translate_sprite(e.dx, e.dy, s0); pshape = translate_shape(e.dx, e.dy, p0);
      return true;
    }
    if (!e.has_sprite) { sprmis++; return false; }
    sprhit++; spr = e.sprite; pshape = e.pshape; return true;
  }
  void addsprite(long id, const Sprite& spr, const Shape& pshape) {  // cache.ml:328-367
    if (!usecache || id < 0) return;
    auto it = tab.find(id);
    if (it != tab.end() && it->second.alias) {
      int dx = it->second.dx, dy = it->second.dy; long t = it->second.target;
      Entry& e = tab[t];
      e.has_sprite = true; e.sprite = translate_sprite(-dx, -dy, spr); e.pshape = translate_shape(-dx, -dy, pshape);
      return;
    }
    Entry& e = tab[id];
    e.has_sprite = true; e.sprite = spr; e.pshape = pshape;
  }
  void addtranslation(long id, long target, int dx, int dy) {  // cache.ml:423-436
    if (!usecache) return;
    auto it = tab.find(target);
    if (it == tab.end()) return;
    Entry e; e.alias = true;
    if (it->second.alias) { e.dx = dx + it->second.dx; e.dy = dy + it->second.dy; e.target = it->second.target; }
    else { e.dx = dx; e.dy = dy; e.target = target; }
    tab[id] = e;
  }
};

struct Renderer {
  Cache cache;
  bool bbox_reject = true;   // render.ml:1270-1279 trivial reject on bounds
  // optional trace of the covered-so-far sets: u after each renderobj at top level
  std::vector<Shape>* trace_u = nullptr;

  // render.ml:469-594 (+ alias translation as served by the cache)
  void shape_of_basicshape(const Obj& o, Shape& shp, Shape& minshp) {
    shape_of_basicshape0(o, shp, minshp);
    if (o.dx || o.dy) { shp = translate_shape(o.dx, o.dy, shp); minshp = translate_shape(o.dx, o.dy, minshp); }
  }
  void shape_of_basicshape0(const Obj& o, Shape& shp, Shape& minshp) {
    switch (o.kind) {
      case Obj::Group: {
        if (cache.getshape(o.id, shp, minshp)) return;
        shp = Shape(); minshp = Shape();
        for (const Obj& c : o.children) {  // members get fresh ids (render.ml:483)
          Shape s, m; shape_of_basicshape(c, s, m);
          shp = shape_union(shp, s);
        }
        cache.addshape(o.id, shp, minshp);  // group minshape is always null (render.ml:494)
        return;
      }
      case Obj::Path: {
        if (cache.getshape(o.id, shp, minshp)) return;
        shapeminshape_of_edgelist(o.edges, o.winding, false, shp, minshp);
        cache.addshape(o.id, shp, minshp);
        return;
      }
      case Obj::Brush: {
        if (cache.getshape(o.id, shp, minshp)) return;
        shp = shape_of_brushstroke(o.stroke); minshp = Shape();  // brush.ml:135-173
        cache.addshape(o.id, shp, minshp);
        return;
      }
      case Obj::Convolved: {  // render.ml:536-555 (cache switched off while computing)
        if (cache.getshape(o.id, shp, minshp)) return;
        int r = radius_of_kernel(o.kernel);
        bool saved = cache.usecache; cache.usecache = false;
        const Obj& child = o.children.at(0);
        Shape cs, cm; shape_of_basicshape(child, cs, cm);
        shp = bloat(r, r, cs);
        if (findfill_fancy(child)) minshp = Shape(); else minshp = erode(r, r, cm);
        cache.usecache = saved;
        cache.addshape(o.id, shp, minshp);
        return;
      }
      case Obj::Filter:  // render.ml:472-474: the filter's geometry
        shape_of_basicshape(o.children.at(0), shp, minshp);
        return;
      case Obj::CPG: {  // render.ml:508-528 (operands are rendered as dummy objects with fresh ids)
        if (cache.getshape(o.id, shp, minshp)) return;
        Shape as, am, bs, bm;
        shape_of_basicshape(o.children.at(0), as, am);
        shape_of_basicshape(o.children.at(1), bs, bm);
        switch (o.cpg_op) {
          case 0: shp = shape_union(as, bs); minshp = shape_union(am, bm); break;
          case 1: shp = shape_intersection(as, bs); minshp = shape_intersection(am, bm); break;
          case 2: shp = shape_difference(as, bm); minshp = shape_difference(am, bs); break;
          default:
            shp = shape_difference(shape_union(as, bs), shape_intersection(am, bm));
            minshp = shape_union(shape_difference(bm, as), shape_difference(am, bs));
        }
        cache.addshape(o.id, shp, minshp);
        return;
      }
      case Obj::Primitive: {
        shp = Shape();
        if (!o.prim_null) shp = shape_box(o.prim[0], o.prim[1], o.prim[2] - o.prim[0] + 1, o.prim[3] - o.prim[1] + 1);
        minshp = shp;
        return;
      }
    }
  }
  static bool findfill_fancy(const Obj& o) {  // render.ml `findfill`
    if (o.kind == Obj::Group) return true;
    if (o.kind == Obj::Convolved) return findfill_fancy(o.children.at(0));
    return o.fill.fancy();
  }

  // polygon.ml:729-746 — AA sprite of a polygon inside `shp`.  Quirk kept: the fill
  // is sampled at the span's first x for every pixel of the span (polygon.ml:736).
  static Sprite polygon_sprite_edgelist(const Fill& f, const Shape& shp, const std::vector<Edge>& edges, Winding w) {
    Shape scaled = mk_scaled_shape(w, edges);
    return map_shape(shp, [&](int x, int y, int l, colour* out) {
      for (int k = 0; k < l; k++) {
        int opacity = pixel_opacity(scaled, x + k, y);
        out[k] = dissolve(f.fillsingle(x, y), opacity);
      }
    });
  }

  // render.ml:984-1078 (+ alias translation as served by the cache)
  Sprite sprite_of_basicshape(const Obj& o, const Shape& shp) {
    if (o.dx || o.dy) return translate_sprite(o.dx, o.dy, sprite_of_basicshape0(o, translate_shape(-o.dx, -o.dy, shp)));
    return sprite_of_basicshape0(o, shp);
  }
  Sprite sprite_of_basicshape0(const Obj& o, const Shape& shp) {
    switch (o.kind) {
      case Obj::Group: {
        Sprite a; Shape u = shp;
        render_scene(u, a, o.children, true);
        return a;
      }
      case Obj::Path: return polygon_sprite_edgelist(o.fill, shp, o.edges, o.sprite_winding);
      case Obj::Brush: return sprite_of_brushstroke(o.stroke, o.fill, shp);
      case Obj::CPG: return sprite_of_cpg(o, shp);
      case Obj::Filter: return sprite_of_basicshape(o.children.at(0), shp);  // render.ml:986-987
      case Obj::Convolved: {  // render.ml:1023-1052: always the "fancy" route
        int r = radius_of_kernel(o.kernel);
        Shape shp2 = bloat(r, r, shp);
        Sprite raster = sprite_of_basicshape(o.children.at(0), shp2);
        return portion(convolve_sprite(o.kernel, raster), shp);
      }
      default: throw std::runtime_error("Internal inconsistency: Should already have been rendered");
    }
  }

  // render.ml:858-864
  static int eor(int a, int b) {
    auto inv = [](int v) { return 255 - v; };
    if (a < 128 && b < 128) return std::max(a, b);
    if (a >= 128 && b < 128) return inv(std::max(inv(a), b));
    if (a < 128 && b >= 128) return inv(std::max(a, inv(b)));
    return std::max(inv(a), inv(b));
  }
  // render.ml:867-981 — constructive planar geometry: both operands are rendered as black alpha
  // mattes, the alphas are combined region by region, then the fill is dissolved by the result.
  Sprite sprite_of_cpg(const Obj& o, const Shape& shp) {
    Obj da = o.children.at(0), db = o.children.at(1);
    da.fill = Fill::plain(mkcol(0, 0, 0)); db.fill = da.fill;
    Shape shp_a, min_a, shp_b, min_b;
    shape_of_basicshape(da, shp_a, min_a); shape_of_basicshape(db, shp_b, min_b);
    shp_a = shape_intersection(shp_a, shp); min_a = shape_intersection(min_a, shp);
    shp_b = shape_intersection(shp_b, shp); min_b = shape_intersection(min_b, shp);
    Shape max_a = shape_difference(shp_a, min_a), max_b = shape_difference(shp_b, min_b);
    Shape tor_a = shape_intersection(shp, shp_a);
    Shape tor_b = shape_difference(shape_intersection(shp, shp_b), shape_intersection(min_a, min_b));
    Sprite spr_a = sprite_of_basicshape(da, tor_a), spr_b = sprite_of_basicshape(db, tor_b);
    Shape rend_a = shape_of_sprite(spr_a), rend_b = shape_of_sprite(spr_b);
    Shape total = shape_union(rend_a, rend_b);
    Shape mm = shape_intersection(shape_intersection(min_a, min_b), total), mM = shape_intersection(shape_intersection(min_a, max_b), total);
    Shape Mm = shape_intersection(shape_intersection(max_a, min_b), total), MM = shape_intersection(shape_intersection(max_a, max_b), total);
    auto invert = [](const Sprite& s) { return sprite_map([](colour c) { return colour_of_channel(255 - alpha_of_colour(c)); }, s); };
    auto both = [&](const CompOp& f) { return caf(f, opaque, portion(spr_a, MM), portion(spr_b, MM)).first; };
    Sprite minmin, minmax, maxmin, maxmax;
    switch (o.cpg_op) {
      case 0:
        minmin = portion(spr_a, mm); minmax = portion(spr_b, mM); maxmin = portion(spr_a, Mm);
        maxmax = both([](colour a, colour b) { int t = alpha_of_colour(a) + alpha_of_colour(b); return colour_of_rgba(0, 0, 0, t > 255 ? 255 : t); });
        break;
      case 2:
        minmax = invert(portion(spr_b, mM));
        maxmax = both([](colour a, colour b) { return colour_of_channel(std::max(0, alpha_of_colour(a) - alpha_of_colour(b))); });
        break;
      case 1:
        minmin = portion(spr_a, mm); minmax = portion(spr_b, mM); maxmin = portion(spr_a, Mm);
        maxmax = both([](colour a, colour b) { return colour_of_channel(std::min(alpha_of_colour(a), alpha_of_colour(b))); });
        break;
      default:
        minmax = invert(portion(spr_b, mM)); maxmin = invert(portion(spr_a, Mm));
        maxmax = both([](colour a, colour b) { return colour_of_rgba(0, 0, 0, eor(alpha_of_colour(a), alpha_of_colour(b))); });
    }
    Shape covered = shape_union(shape_union(mm, mM), shape_union(Mm, MM));
    Sprite parts[8] = {minmin, minmax, maxmin, maxmax,
                       portion(spr_a, shape_intersection(shape_difference(min_a, covered), rend_a)),
                       portion(spr_b, shape_intersection(shape_difference(min_b, covered), rend_b)),
                       portion(spr_a, shape_intersection(shape_difference(max_a, covered), rend_a)),
                       portion(spr_b, shape_intersection(shape_difference(max_b, covered), rend_b))};
    Sprite alpha;
    for (auto& part : parts) alpha = caf([](colour, colour) -> colour { throw std::runtime_error("CPG caf"); }, opaque, alpha, part).first;
    // 6. apply the fill (map_coords: per pixel fill lookup)
    Sprite out = alpha;
    for (auto& r : out.rows) {
      int off = 0;
      for (auto& sp : r.spans) { for (int k = 0; k < sp.len; k++) r.px[off + k] = dissolve(o.fill.fillsingle(sp.x + k, r.y), alpha_of_colour(r.px[off + k])); off += sp.len; }
    }
    return out;
  }

  // render.ml:1134-1242 (non-filter objects)
  Sprite spriteof(const Obj& o, const Shape& shp) {
    Sprite cached; Shape pshape;
    if (cache.getsprite(o.id, cached, pshape) && (o.dx || o.dy)) {  // entries hold the untranslated geometry's sprite
      cached = translate_sprite(o.dx, o.dy, cached); pshape = translate_shape(o.dx, o.dy, pshape);
    }
    Shape shptorender = shape_difference(shp, pshape);
    if (shptorender.null()) return portion(cached, shp);
    Sprite rendered;
    if (o.kind == Obj::Primitive) {
      Shape s, m; shape_of_basicshape(o, s, m);
      rendered = fillshape(shape_intersection(shptorender, s), Fill::plain(o.prim_colour));
    } else {
      Shape s, m; shape_of_basicshape(o, s, m);
      Shape maxshape = shape_difference(s, m);
      Sprite maxbit = sprite_of_basicshape(o, shape_intersection(shptorender, maxshape));
      Sprite minbit = translate_sprite(o.dx, o.dy, fillshape(translate_shape(-o.dx, -o.dy, shape_intersection(m, shptorender)), o.fill));
      rendered = caf(nocover, opaque, minbit, maxbit).first;
    }
    Sprite newwhole = caf(nocover, opaque, cached, rendered).first;
    Shape pshape2 = shape_of_sprite(newwhole);
    if (o.kind != Obj::Primitive) cache.addsprite(o.id, translate_sprite(-o.dx, -o.dy, newwhole), translate_shape(-o.dx, -o.dy, pshape2));
    return portion(newwhole, shape_intersection(shp, pshape2));
  }

  // render.ml:1080-1131 — a filter object: render the reading scene, filter it, render the scene
  // below where the filter's matte is not opaque, blend the two by the matte (blend', 1248-1265).
  // Returns the sprite and the extra finish (the whole shape of the filter's geometry).
  std::pair<Sprite, Shape> spriteof_filter(const Obj& o, const Scene& objs, size_t below, const Shape& shptorender) {
    Scene tail(objs.begin() + (long)below, objs.end());
    int r = o.filter_kind == 3 ? radius_of_kernel(o.kernel) : 0;
    // reading_scene (filters.ml:221, 233-234, 247-250, 276-278)
    Shape readshape = o.filter_kind == 3 ? bloat(2 * r + 1, 2 * r + 1, shptorender) : shptorender;
    Shape shp2 = shptorender;
    Scene none, tl;
    if (o.filter_kind == 5) {  // Filters.minus (filters.ml:289-303): hd scene is cut away inside the filter
      if (tail.empty()) throw std::runtime_error("hd");
      Shape fs, fm, hs, hm; shape_of_basicshape(o, fs, fm); shape_of_basicshape(tail[0], hs, hm);
      shp2 = shape_intersection(shape_intersection(fs, hs), shptorender);
      readshape = shp2;
      tl = Scene(tail.begin() + 1, tail.end());
    }
    if (o.filter_kind == 6) {  // Filters.smear (filters.ml:201-217): read in bloat rx ry shp
      const int rr = (o.stroke.bw() - 1) / 2;
      readshape = bloat(rr, rr, shptorender);
    }
    const Scene& scene2 = o.filter_kind == 1 ? none : (o.filter_kind == 4 ? *o.reading_scene : (o.filter_kind == 5 ? tl : tail));
    Sprite X; { Shape u = readshape; render_scene(u, X, scene2, true); }
    Sprite Y;
    switch (o.filter_kind) {
      case 2: Y = sprite_map(monochrome, X); break;                                  // filters.ml:235-236
      case 3: {                                                                        // filters.ml:251-255
        Shape bloated = bloat(r, r, shape_of_sprite(X));
        Y = convolve_sprite_in_shape(o.kernel, X, bloated, shape_intersection(bloated, shp2));
        break;
      }
      case 6: {                                                                        // filters.ml:211-214
        Sprite sm = smear(X, o.stroke, o.smear_pts);
        Y = portion(sm, shape_intersection(shp2, shape_of_sprite(sm)));
        break;
      }
      default: Y = X;                                                                  // nullfilterfunction
    }
    Sprite alpha = sprite_of_basicshape(o, shp2);
    Shape finished = caf(nocover, opaque, Sprite(), alpha).second;
    Sprite Z; { Shape u = shape_difference(shp2, finished); if (!u.null()) render_scene(u, Z, tail, true); }
    // blend' (render.ml:1248-1265)
    Sprite a_in_z = portion(alpha, shape_of_sprite(Z)), a_in_y = portion(alpha, shape_of_sprite(Y));
    Sprite z_att = caf([](colour c, colour al) { return dissolve(c, 255 - alpha_of_colour(al)); }, opaque, Z, a_in_z).first;
    Sprite y_att = caf([](colour c, colour al) { return dissolve(c, alpha_of_colour(al)); }, opaque, Y, a_in_y).first;
    Sprite res = caf(pd_plus, opaque, z_att, y_att).first;
    Shape s, m; shape_of_basicshape(o, s, m);
    return {res, s};
  }

  // render.ml:1268-1308
  void renderobj(const Scene& objs, size_t idx, Shape& u, Sprite& a) {
    const Obj& o = objs[idx];
    if (bbox_reject && o.has_bounds) {
      Box ub; shape_bounds(u, ub);
      // Pdfutil.box_overlap on inclusive integer boxes
      if (o.bounds[0] + o.dx > ub.x1 || o.bounds[1] + o.dx < ub.x0 || o.bounds[2] + o.dy > ub.y1 || o.bounds[3] + o.dy < ub.y0) return;
    }
    Shape r, rm; shape_of_basicshape(o, r, rm);
    Shape r2 = shape_intersection(r, u);
    if (r2.null()) return;
    if (o.kind == Obj::Filter) {  // render.ml:1289-1291, 1308: composited with over; only the extra finish leaves u
      auto se = spriteof_filter(o, objs, idx + 1, r2);
      a = caf(over, opaque, a, se.first).first;
      u = shape_difference(u, se.second);
      return;
    }
    Sprite s = spriteof(o, r2);
    if (o.pretrans >= 0) {
      int d = o.pretrans;
      s = sprite_map([d](colour c) { return dissolve(c, d); }, s);
    }
    auto res = caf(over, opaque, a, s);
    a = std::move(res.first);
    u = shape_difference(u, res.second);
  }
  // render.ml:1310-1335
  void render_scene(Shape& u, Sprite& a, const Scene& objs, bool nested) {
    for (size_t i = 0; i < objs.size(); i++) {
      if (u.null()) return;
      renderobj(objs, i, u, a);
      if (!nested && trace_u) trace_u->push_back(u);
    }
  }
  // render.ml:1345-1365 — scene pass and background pass over the same update.
  Sprite render_frame(const Scene& scene, const Scene& background, const Shape& update) {
    Shape u1 = update; Sprite a1; render_scene(u1, a1, scene, false);
    Shape u2 = update; Sprite a2; render_scene(u2, a2, background, true);
    return caf(over, opaque, a1, a2).first;
  }
  // render.ml:1368-1370
  Sprite render_simple_scene(const Scene& scene, const Shape& shape) {
    Shape u = shape; Sprite a; render_scene(u, a, scene, false);
    return a;
  }
};

// render.ml:1376-1400 — dirty region of a moved object (o = old, n = new shapes).
inline Shape plaindirty(const Shape& shp_o, const Shape& min_o, const Shape& shp_n, const Shape& min_n, const Shape& u) {
  return shape_intersection(shape_union(shape_difference(shp_o, min_n), shape_difference(shp_n, min_o)), u);
}
inline Shape alldirty(const Shape& shp_o, const Shape& shp_n, const Shape& u) {
  return shape_intersection(shape_union(shp_o, shp_n), u);
}
}  // namespace oracle
