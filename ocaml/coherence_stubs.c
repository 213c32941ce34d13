/* coherence_stubs.c — OCaml externals over the C ABI (include/coherence_b200.h): one stub per exported symbol.
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no OCaml toolchain here (no ocamlfind,
 * no caml/mlvalues.h; probed on the GPU box as well).  Delivered as reviewed source; build with
 *   ocamlfind ocamlopt -package bigarray -I include -c ocaml/coherence_stubs.c
 * and link with -cclib -lcoherence_b200.
 *
 * Conventions: bulk data arrives in Bigarray.Array1 (c_layout, int32 / int8_unsigned / float64); handles
 * are boxed int64 (Int64.t); a non-zero status raises Failure (coh_last_error), which is the
 * reference's own error convention (failwith).  The stubs never retain OCaml pointers across
 * calls, so the GC may move values freely; device objects are freed by explicit calls wired to
 * Gc.finalise in coherence_gpu.ml.  The caller is the single OCaml thread; calls block until the
 * requested host bytes are ready, so the runtime lock is not released.
 *
 * GC discipline: every value that is allocated here is bound to a CAMLlocal root BEFORE it is stored into another
 * block (Store_field (t, i, caml_copy_int64 (..)) would compute the field address before the allocation can move
 * the tuple); Bigarray arguments are checked against the sizes the C ABI is going to write.
 */
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <string.h>
#include "coherence_b200.h"

#define CTX(v) ((coh_ctx*)Nativeint_val(v))
#define SHAPE(v) ((coh_shape_t)Int64_val(v))
#define SCENE(v) ((coh_scene_t)Int64_val(v))
#define BA_LEN(v) ((int64_t)Caml_ba_array_val(v)->dim[0])
static void check(coh_ctx* c, int rc) { if (rc) caml_failwith(coh_last_error(c)); }
static void need(value ba, int64_t n, const char* what) { if (BA_LEN(ba) < n) caml_invalid_argument(what); }
/* (int64, int64) tuple of two handles */
static value pair_of_handles(coh_shape_t a, coh_shape_t b) {
  CAMLparam0();
  CAMLlocal3(pair, va, vb);
  va = caml_copy_int64((int64_t)a);
  vb = caml_copy_int64((int64_t)b);
  pair = caml_alloc_tuple(2);
  Store_field(pair, 0, va);
  Store_field(pair, 1, vb);
  CAMLreturn(pair);
}
static value tuple_of_ints(const int32_t* v, int n) {
  CAMLparam0();
  CAMLlocal1(r);
  r = caml_alloc_tuple(n);
  for (int k = 0; k < n; k++) Store_field(r, k, Val_int(v[k]));   /* immediates: no allocation */
  CAMLreturn(r);
}

/* ---- lifecycle ---- */
CAMLprim value coh_ml_init(value device) {
  CAMLparam1(device);
  coh_ctx* c = NULL;
  if (coh_init(Int_val(device), &c)) caml_failwith(coh_last_error(NULL));
  CAMLreturn(caml_copy_nativeint((intnat)c));
}
CAMLprim value coh_ml_shutdown(value ctx) { coh_shutdown(CTX(ctx)); return Val_unit; }
CAMLprim value coh_ml_device_name(value ctx) {
  CAMLparam1(ctx);
  char buf[256];
  check(CTX(ctx), coh_device_name(CTX(ctx), buf, sizeof buf));
  CAMLreturn(caml_copy_string(buf));
}
CAMLprim value coh_ml_stream(value ctx) { return caml_copy_nativeint((intnat)coh_stream(CTX(ctx))); }
CAMLprim value coh_ml_set_stream(value ctx, value s) { check(CTX(ctx), coh_set_stream(CTX(ctx), (void*)Nativeint_val(s))); return Val_unit; }
CAMLprim value coh_ml_launch_count(value ctx) { return caml_copy_int64(coh_launch_count(CTX(ctx))); }
CAMLprim value coh_ml_set_timing(value ctx, value on) { check(CTX(ctx), coh_set_timing(CTX(ctx), Bool_val(on))); return Val_unit; }
CAMLprim value coh_ml_get_timing(value ctx) {   /* (raster ms, binning ms, frames) */
  CAMLparam1(ctx);
  CAMLlocal3(r, a, b);
  double w = 0, bn = 0; int64_t n = 0;
  check(CTX(ctx), coh_get_timing(CTX(ctx), &w, &bn, &n));
  a = caml_copy_double(w); b = caml_copy_double(bn);
  r = caml_alloc_tuple(3);
  Store_field(r, 0, a); Store_field(r, 1, b); Store_field(r, 2, Val_long((long)n));
  CAMLreturn(r);
}
CAMLprim value coh_ml_set_option(value ctx, value name, value v) { check(CTX(ctx), coh_set_option(CTX(ctx), String_val(name), Int_val(v))); return Val_unit; }
CAMLprim value coh_ml_mem_in_use(value ctx) { int64_t n = 0; check(CTX(ctx), coh_mem_in_use(CTX(ctx), &n)); return caml_copy_int64(n); }
CAMLprim value coh_ml_sync(value ctx) { check(CTX(ctx), coh_sync(CTX(ctx))); return Val_unit; }

/* ---- Colour.colour (31-bit int) <-> RGBA8 word ---- */
CAMLprim value coh_ml_rgba8_of_colour(value c) { return caml_copy_int32((int32_t)coh_rgba8_of_colour((int32_t)Long_val(c))); }
CAMLprim value coh_ml_colour_of_rgba8(value w) { return Val_long(coh_colour_of_rgba8((uint32_t)Int32_val(w))); }

/* ---- Polygon ---- */
/* Polygon.shapeminshape_of_unsorted_edgelist: edges : (int32, c_layout) Array1 of 4*n */
CAMLprim value coh_ml_shapeminshape(value ctx, value edges, value winding) {
  CAMLparam3(ctx, edges, winding);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_shapeminshape_of_edgelist(CTX(ctx), (const int32_t*)Caml_ba_data_val(edges), (int32_t)(BA_LEN(edges) / 4), Int_val(winding), &s, &m));
  CAMLreturn(pair_of_handles(s, m));
}
/* ---- brush strokes outside a scene (brush.mli:20-27); brush = one packed BRUSH coh_object record (coh_ml_pack_object) ---- */
CAMLprim value coh_ml_brush_shape(value ctx, value brush, value points) {
  CAMLparam3(ctx, brush, points);
  coh_shape_t s = 0;
  check(CTX(ctx), coh_brush_shape(CTX(ctx), (const coh_object*)Caml_ba_data_val(brush), (const int32_t*)Caml_ba_data_val(points), (int32_t)(BA_LEN(points) / 2), &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_brush_sprite(value ctx, value brush, value points, value shape, value out) {
  CAMLparam5(ctx, brush, points, shape, out);
  int64_t n = 0;
  check(CTX(ctx), coh_brush_sprite(CTX(ctx), (const coh_object*)Caml_ba_data_val(brush), (const int32_t*)Caml_ba_data_val(points), (int32_t)(BA_LEN(points) / 2), SHAPE(shape),
        (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}
/* returns (handle of the result's shape, number of pixels written to out) */
CAMLprim value coh_ml_brush_smear(value ctx, value shape, value rgba, value brush, value points, value smear_points, value out) {
  CAMLparam5(ctx, shape, rgba, brush, points);
  CAMLxparam2(smear_points, out);
  CAMLlocal2(pair, h);
  coh_shape_t so = 0; int64_t n = 0;
  check(CTX(ctx), coh_brush_smear(CTX(ctx), SHAPE(shape), (const uint32_t*)Caml_ba_data_val(rgba), (const coh_object*)Caml_ba_data_val(brush),
        (const int32_t*)Caml_ba_data_val(points), (int32_t)(BA_LEN(points) / 2), (const int32_t*)Caml_ba_data_val(smear_points), (int32_t)(BA_LEN(smear_points) / 2),
        &so, (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  h = caml_copy_int64((int64_t)so);
  pair = caml_alloc_tuple(2);
  Store_field(pair, 0, h); Store_field(pair, 1, Val_long((long)n));
  CAMLreturn(pair);
}
CAMLprim value coh_ml_brush_smear_bc(value* argv, int argn) { (void)argn; return coh_ml_brush_smear(argv[0], argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]); }
/* N2: Polygon.edgelist_of_path on the device; segs : float64 Array1 of 9*n, out : int32 Array1 of 4*cap; returns the edge count */
CAMLprim value coh_ml_edgelist_of_path(value ctx, value segs, value out) {
  CAMLparam3(ctx, segs, out);
  int64_t n = 0;
  check(CTX(ctx), coh_edgelist_of_path(CTX(ctx), (const double*)Caml_ba_data_val(segs), (int32_t)(BA_LEN(segs) / 9), (int32_t*)Caml_ba_data_val(out), BA_LEN(out) / 4, &n));
  CAMLreturn(Val_long((long)n));
}
/* Polygon.shapeminshape_polygon with the flattened edges kept in HBM */
CAMLprim value coh_ml_shapeminshape_of_path(value ctx, value segs, value winding) {
  CAMLparam3(ctx, segs, winding);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_shapeminshape_of_path(CTX(ctx), (const double*)Caml_ba_data_val(segs), (int32_t)(BA_LEN(segs) / 9), Int_val(winding), &s, &m));
  CAMLreturn(pair_of_handles(s, m));
}
/* N2: Shapes.strokepath.  spec : float64 Array1 [| startcap; join; endcap; mitrelimit; linewidth |] (caps Butt 0 / Round 1 /
 * Projecting 2, joins Round 0 / Mitred 1 / Bevel 2); counts : int32 Array1, the segments of each subpath. */
static coh_strokespec strokespec_of(value spec) {
  const double* d = (const double*)Caml_ba_data_val(spec);
  coh_strokespec sp; memset(&sp, 0, sizeof sp);
  need(spec, 5, "strokespec: startcap, join, endcap, mitrelimit, linewidth");
  sp.startcap = (int32_t)d[0]; sp.join = (int32_t)d[1]; sp.endcap = (int32_t)d[2]; sp.mitrelimit = d[3]; sp.linewidth = d[4];
  return sp;
}
/* returns (edge count, winding rule of the outline); out : int32 Array1 of 4 * cap */
CAMLprim value coh_ml_strokepath(value ctx, value spec, value segs, value counts, value out) {
  CAMLparam5(ctx, spec, segs, counts, out);
  CAMLlocal1(r);
  coh_strokespec sp = strokespec_of(spec);
  int64_t n = 0; int32_t w = 0;
  check(CTX(ctx), coh_strokepath(CTX(ctx), &sp, (const double*)Caml_ba_data_val(segs), (const int32_t*)Caml_ba_data_val(counts), (int32_t)BA_LEN(counts),
                                 (int32_t*)Caml_ba_data_val(out), BA_LEN(out) / 4, &n, &w));
  r = caml_alloc_tuple(2);
  Store_field(r, 0, Val_long((long)n)); Store_field(r, 1, Val_int(w));
  CAMLreturn(r);
}
CAMLprim value coh_ml_shapeminshape_of_stroke(value ctx, value spec, value segs, value counts) {
  CAMLparam4(ctx, spec, segs, counts);
  coh_strokespec sp = strokespec_of(spec);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_shapeminshape_of_stroke(CTX(ctx), &sp, (const double*)Caml_ba_data_val(segs), (const int32_t*)Caml_ba_data_val(counts), (int32_t)BA_LEN(counts), &s, &m));
  CAMLreturn(pair_of_handles(s, m));
}
/* Shapes.bounds_stroke: (xmin, xmax, ymin, ymax) */
CAMLprim value coh_ml_host_bounds_stroke(value spec, value segs, value counts) {
  CAMLparam3(spec, segs, counts);
  CAMLlocal1(r);
  coh_strokespec sp = strokespec_of(spec);
  int32_t b[4];
  if (coh_host_bounds_stroke(&sp, (const double*)Caml_ba_data_val(segs), (const int32_t*)Caml_ba_data_val(counts), (int32_t)BA_LEN(counts), b) != 0)
    caml_failwith("Polygon2.bounds_polygon: Malformed (empty) path");
  r = caml_alloc_tuple(4);
  for (int k = 0; k < 4; k++) Store_field(r, k, Val_int(b[k]));
  CAMLreturn(r);
}
/* Shapes.strokepath_polygon on the host: returns (outline segments, outline subpaths, winding); out : float64 Array1 of
 * 9 * cap, out_counts : int32 Array1 */
CAMLprim value coh_ml_host_strokepath(value spec, value segs, value counts, value out, value out_counts) {
  CAMLparam5(spec, segs, counts, out, out_counts);
  CAMLlocal1(r);
  coh_strokespec sp = strokespec_of(spec);
  int32_t m = 0, w = 0;
  int64_t n = coh_host_strokepath(&sp, (const double*)Caml_ba_data_val(segs), (const int32_t*)Caml_ba_data_val(counts), (int32_t)BA_LEN(counts),
                                  (double*)Caml_ba_data_val(out), BA_LEN(out) / 9, (int32_t*)Caml_ba_data_val(out_counts), (int32_t)BA_LEN(out_counts), &m, &w);
  if (n < 0) caml_failwith("Shapes.joinsegments: Not implemented");
  r = caml_alloc_tuple(3);
  Store_field(r, 0, Val_long((long)n)); Store_field(r, 1, Val_int(m)); Store_field(r, 2, Val_int(w));
  CAMLreturn(r);
}
/* opacity bytes of every pixel of `shape`, span order; out : (int, int8_unsigned) Array1 of Sprite.shape_card shape */
CAMLprim value coh_ml_polygon_opacity(value ctx, value edges, value winding, value shape, value out) {
  CAMLparam5(ctx, edges, winding, shape, out);
  int64_t n = 0;
  check(CTX(ctx), coh_polygon_opacity(CTX(ctx), (const int32_t*)Caml_ba_data_val(edges), (int32_t)(BA_LEN(edges) / 4), Int_val(winding), SHAPE(shape),
        (uint8_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}
/* Polygon.polygon_sprite_edgelist: fill = one packed coh_object record (coh_ml_pack_object); out : int32 Array1 */
CAMLprim value coh_ml_polygon_sprite(value ctx, value fill, value edges, value winding, value shape, value out) {
  CAMLparam5(ctx, fill, edges, winding, shape);
  CAMLxparam1(out);
  int64_t n = 0;
  need(fill, (int64_t)sizeof(coh_object), "coh_polygon_sprite: fill record too short");
  check(CTX(ctx), coh_polygon_sprite(CTX(ctx), (const coh_object*)Caml_ba_data_val(fill), (const int32_t*)Caml_ba_data_val(edges), (int32_t)(BA_LEN(edges) / 4),
        Int_val(winding), SHAPE(shape), (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}
CAMLprim value coh_ml_polygon_sprite_bc(value* a, int n) { (void)n; return coh_ml_polygon_sprite(a[0], a[1], a[2], a[3], a[4], a[5]); }

/* ---- Sprite.shape ---- */
CAMLprim value coh_ml_shape_box(value ctx, value x, value y, value w, value h) {
  CAMLparam5(ctx, x, y, w, h);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_shape_box(CTX(ctx), Int_val(x), Int_val(y), Int_val(w), Int_val(h), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}
/* Sprite.shape export: returns the flat (y, nspans, (x,len)...) records in a fresh Bigarray */
CAMLprim value coh_ml_shape_export(value ctx, value shape) {
  CAMLparam2(ctx, shape);
  CAMLlocal1(ba);
  int64_t n = 0, got = 0;
  check(CTX(ctx), coh_shape_export_size(CTX(ctx), SHAPE(shape), &n));
  intnat dim = (intnat)n;
  ba = caml_ba_alloc(CAML_BA_INT32 | CAML_BA_C_LAYOUT, 1, NULL, &dim);
  check(CTX(ctx), coh_shape_export(CTX(ctx), SHAPE(shape), (int32_t*)Caml_ba_data_val(ba), n, &got));
  CAMLreturn(ba);
}
CAMLprim value coh_ml_shape_import(value ctx, value flat) {
  CAMLparam2(ctx, flat);
  coh_shape_t s = 0;
  check(CTX(ctx), coh_shape_import(CTX(ctx), (const int32_t*)Caml_ba_data_val(flat), BA_LEN(flat), &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_shape_bounds(value ctx, value s) {   /* (x0, y0, x1, y1) option */
  CAMLparam2(ctx, s);
  CAMLlocal2(t, some);
  int32_t box[4], isnull = 0;
  check(CTX(ctx), coh_shape_bounds(CTX(ctx), SHAPE(s), box, &isnull));
  if (isnull) CAMLreturn(Val_int(0));   /* None */
  t = tuple_of_ints(box, 4);
  some = caml_alloc(1, 0);
  Store_field(some, 0, t);
  CAMLreturn(some);
}
CAMLprim value coh_ml_shape_card(value ctx, value s) { int64_t n = 0; check(CTX(ctx), coh_shape_card(CTX(ctx), SHAPE(s), &n)); return Val_long((long)n); }
CAMLprim value coh_ml_shape_free(value ctx, value s) { coh_shape_free(CTX(ctx), SHAPE(s)); return Val_unit; }
#define BINOP(name, fn)                                                              \
  CAMLprim value name(value ctx, value a, value b) {                                 \
    CAMLparam3(ctx, a, b);                                                           \
    coh_shape_t o = 0;                                                               \
    check(CTX(ctx), fn(CTX(ctx), SHAPE(a), SHAPE(b), &o));                           \
    CAMLreturn(caml_copy_int64((int64_t)o));                                         \
  }
BINOP(coh_ml_shape_union, coh_shape_union)              /* Sprite.( ||| ) */
BINOP(coh_ml_shape_difference, coh_shape_difference)    /* Sprite.( --- ) */
BINOP(coh_ml_shape_intersection, coh_shape_intersection)/* Sprite.( &&& ) */
#define UNOP(name, fn)                                                               \
  CAMLprim value name(value ctx, value a, value m, value n) {                        \
    CAMLparam4(ctx, a, m, n);                                                        \
    coh_shape_t o = 0;                                                               \
    check(CTX(ctx), fn(CTX(ctx), SHAPE(a), Int_val(m), Int_val(n), &o));             \
    CAMLreturn(caml_copy_int64((int64_t)o));                                         \
  }
UNOP(coh_ml_shape_translate, coh_shape_translate)       /* Sprite.translate_shape */
UNOP(coh_ml_shape_bloat, coh_shape_bloat)               /* Sprite.bloat */
UNOP(coh_ml_shape_erode, coh_shape_erode)               /* Sprite.erode */

/* ---- whole sprites: (shape handle, RGBA8 per pixel in span order as an int32 Bigarray) ---- */
CAMLprim value coh_ml_shape_intersects(value ctx, value a, value b) {
  int32_t yes = 0;
  check(CTX(ctx), coh_shape_intersects(CTX(ctx), SHAPE(a), SHAPE(b), &yes));
  return Val_bool(yes);
}
CAMLprim value coh_ml_sprite_portion(value ctx, value shape, value rgba, value sub, value out) {
  CAMLparam5(ctx, shape, rgba, sub, out);
  int64_t n = 0;
  check(CTX(ctx), coh_sprite_portion(CTX(ctx), SHAPE(shape), (const uint32_t*)Caml_ba_data_val(rgba), SHAPE(sub), (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}
CAMLprim value coh_ml_sprite_fillshape(value ctx, value shape, value fill, value out) {
  CAMLparam4(ctx, shape, fill, out);
  int64_t n = 0;
  need(fill, (int64_t)sizeof(coh_object), "sprite_fillshape: fill record too short");
  check(CTX(ctx), coh_sprite_fillshape(CTX(ctx), SHAPE(shape), (const coh_object*)Caml_ba_data_val(fill), (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}
CAMLprim value coh_ml_sprite_map(value ctx, value op_arg, value rgba, value out) {
  CAMLparam4(ctx, op_arg, rgba, out);
  need(out, BA_LEN(rgba), "sprite_map: output shorter than input");
  check(CTX(ctx), coh_sprite_map(CTX(ctx), Int_val(Field(op_arg, 0)), Int_val(Field(op_arg, 1)), (const uint32_t*)Caml_ba_data_val(rgba), BA_LEN(rgba), (uint32_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_sprite_map_coords_fill(value ctx, value shape, value fill, value rgba, value out) {
  CAMLparam5(ctx, shape, fill, rgba, out);
  int64_t n = 0;
  need(fill, (int64_t)sizeof(coh_object), "sprite_map_coords_fill: fill record too short");
  check(CTX(ctx), coh_sprite_map_coords_fill(CTX(ctx), SHAPE(shape), (const coh_object*)Caml_ba_data_val(fill), (const uint32_t*)Caml_ba_data_val(rgba), (uint32_t*)Caml_ba_data_val(out), BA_LEN(out), &n));
  CAMLreturn(Val_long((long)n));
}

/* ---- Convolve.convolve_sprite kernel sprite: sprite = (shape handle, RGBA8 per pixel in span order) ---- */
CAMLprim value coh_ml_convolve_sprite(value ctx, value kind_r, value shape, value rgba_in, value rgba_out) {
  CAMLparam5(ctx, kind_r, shape, rgba_in, rgba_out);
  coh_shape_t o = 0;
  int64_t n = 0;
  check(CTX(ctx), coh_convolve_sprite(CTX(ctx), Int_val(Field(kind_r, 0)), Int_val(Field(kind_r, 1)), SHAPE(shape),
        (const uint32_t*)Caml_ba_data_val(rgba_in), &o, (uint32_t*)Caml_ba_data_val(rgba_out), BA_LEN(rgba_out), &n));
  CAMLreturn(caml_copy_int64((int64_t)o));
}

/* ---- Cache ---- */
CAMLprim value coh_ml_cache_configure(value ctx, value on, value bytes) {   /* Cache.usecache / Cache.setsize */
  check(CTX(ctx), coh_cache_configure(CTX(ctx), Bool_val(on), Int64_val(bytes)));
  return Val_unit;
}
CAMLprim value coh_ml_cache_clear(value ctx) { check(CTX(ctx), coh_cache_clear(CTX(ctx))); return Val_unit; }
CAMLprim value coh_ml_cache_stats(value ctx) {   /* (hits, misses, bytes, entries, sprite hits, sprite bytes) */
  CAMLparam1(ctx);
  CAMLlocal1(r);
  int64_t st[6] = {0, 0, 0, 0, 0, 0};
  check(CTX(ctx), coh_cache_stats(CTX(ctx), st));
  r = caml_alloc_tuple(4);
  for (int k = 0; k < 4; k++) Store_field(r, k, Val_long((long)st[k]));
  CAMLreturn(r);
}
CAMLprim value coh_ml_cache_sprite_stats(value ctx, value scene) {   /* (frames from sprites, frames that filled, bytes, objects) */
  CAMLparam2(ctx, scene);
  int64_t st[4] = {0, 0, 0, 0};
  int32_t v[4];
  check(CTX(ctx), coh_cache_sprite_stats(CTX(ctx), SCENE(scene), st));
  for (int k = 0; k < 4; k++) v[k] = (int32_t)st[k];
  CAMLreturn(tuple_of_ints(v, 4));
}
CAMLprim value coh_ml_cache_addshape(value ctx, value id, value s, value m) {
  check(CTX(ctx), coh_cache_addshape(CTX(ctx), Int64_val(id), SHAPE(s), SHAPE(m)));
  return Val_unit;
}
CAMLprim value coh_ml_cache_getshape(value ctx, value id) {   /* (shape_h * shape_h) option */
  CAMLparam2(ctx, id);
  CAMLlocal2(p, some);
  coh_shape_t s = 0, m = 0; int32_t found = 0;
  check(CTX(ctx), coh_cache_getshape(CTX(ctx), Int64_val(id), &s, &m, &found));
  if (!found) CAMLreturn(Val_int(0));
  p = pair_of_handles(s, m);
  some = caml_alloc(1, 0);
  Store_field(some, 0, p);
  CAMLreturn(some);
}
CAMLprim value coh_ml_cache_addtranslation(value ctx, value id, value target, value dx, value dy) {
  check(CTX(ctx), coh_cache_addtranslation(CTX(ctx), Int64_val(id), Int64_val(target), Int_val(dx), Int_val(dy)));
  return Val_unit;
}
/* Render.plaindirty (plain = true) / alldirty */
CAMLprim value coh_ml_dirty_region(value ctx, value shapes, value u, value plain) {
  CAMLparam4(ctx, shapes, u, plain);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_dirty_region(CTX(ctx), SHAPE(Field(shapes, 0)), SHAPE(Field(shapes, 1)), SHAPE(Field(shapes, 2)), SHAPE(Field(shapes, 3)),
        SHAPE(u), Bool_val(plain), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}

/* ---- Render ---- */
/* One coh_object record from its fields, in the order of the header: ints = [| kind; winding; first; count; fill_kind;
 * colour0; colour1; fill_flags; pretrans; dx; dy; b0; b1; b2; b3; p0; p1; p2; p3; prim_null; convolve; sprite_winding;
 * first2; count2; winding2; cpg_op; filter_kind; filter_kernel |], floats = [| fparam (6); brush_opacity; brush_radius |].
 * The struct layout stays on the C side. */
CAMLprim value coh_ml_pack_object(value objs, value index, value ints, value floats, value id) {
  CAMLparam5(objs, index, ints, floats, id);
  const int64_t i = Long_val(index);
  need(objs, (i + 1) * (int64_t)sizeof(coh_object), "coh_ml_pack_object: record outside the buffer");
  if (Wosize_val(ints) < 28 || Wosize_val(floats) / Double_wosize < 8) caml_invalid_argument("coh_ml_pack_object: field arrays too short");
  coh_object o;
  memset(&o, 0, sizeof o);
#define I(k) ((int32_t)Long_val(Field(ints, k)))
  o.kind = I(0); o.winding = I(1); o.first = I(2); o.count = I(3); o.fill_kind = I(4);
  o.colour0 = (uint32_t)I(5); o.colour1 = (uint32_t)I(6); o.fill_flags = I(7); o.pretrans = I(8); o.dx = I(9); o.dy = I(10);
  for (int k = 0; k < 4; k++) { o.bounds[k] = I(11 + k); o.prim[k] = I(15 + k); }
  o.prim_null = I(19); o.convolve = I(20); o.sprite_winding = I(21);
  o.first2 = I(22); o.count2 = I(23); o.winding2 = I(24); o.cpg_op = I(25); o.filter_kind = I(26); o.filter_kernel = I(27);
#undef I
  for (int k = 0; k < 6; k++) o.fparam[k] = Double_flat_field(floats, k);
  o.brush_opacity = Double_flat_field(floats, 6); o.brush_radius = Double_flat_field(floats, 7);
  o.id = Int64_val(id);
  memcpy((char*)Caml_ba_data_val(objs) + i * sizeof(coh_object), &o, sizeof o);
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_sizeof_object(value unit) { (void)unit; return Val_long((long)sizeof(coh_object)); }
/* objs is a Bigarray of bytes holding n packed coh_object records built by Coherence_gpu.flatten_scene */
CAMLprim value coh_ml_scene_create(value ctx, value objs, value n_background, value edges, value points) {
  CAMLparam5(ctx, objs, n_background, edges, points);
  coh_scene_t s = 0;
  check(CTX(ctx), coh_scene_create(CTX(ctx), (const coh_object*)Caml_ba_data_val(objs), (int32_t)(BA_LEN(objs) / (int64_t)sizeof(coh_object)), Int_val(n_background),
        (const int32_t*)Caml_ba_data_val(edges), (int32_t)(BA_LEN(edges) / 4), (const int32_t*)Caml_ba_data_val(points), (int32_t)(BA_LEN(points) / 2), &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_scene_free(value ctx, value s) { coh_scene_free(CTX(ctx), SCENE(s)); return Val_unit; }
CAMLprim value coh_ml_fb_configure(value ctx, value w, value h, value y0, value y1) {
  check(CTX(ctx), coh_fb_configure(CTX(ctx), Int_val(w), Int_val(h), Int_val(y0), Int_val(y1)));
  return Val_unit;
}
CAMLprim value coh_ml_fb_attach(value ctx, value p) { check(CTX(ctx), coh_fb_attach(CTX(ctx), (void*)Nativeint_val(p))); return Val_unit; }
CAMLprim value coh_ml_fb_device_ptr(value ctx) { return caml_copy_nativeint((intnat)coh_fb_device_ptr(CTX(ctx))); }
CAMLprim value coh_ml_fb_set_peers(value ctx, value peers) {   /* nativeint array */
  CAMLparam2(ctx, peers);
  void* p[8];
  const int n = (int)Wosize_val(peers);
  if (n > 7) caml_invalid_argument("coh_fb_set_peers: at most 7 peers");
  for (int k = 0; k < n; k++) p[k] = (void*)Nativeint_val(Field(peers, k));
  check(CTX(ctx), coh_fb_set_peers(CTX(ctx), n, p));
  CAMLreturn(Val_unit);
}
/* Render.render_frame over update = Sprite.box x y w h (flags: 1 = keep the covered-so-far set) */
CAMLprim value coh_ml_render_frame(value ctx, value scene, value box, value flags) {
  CAMLparam4(ctx, scene, box, flags);
  check(CTX(ctx), coh_render_frame(CTX(ctx), SCENE(scene), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), Int_val(flags)));
  CAMLreturn(Val_unit);
}
/* ... then plot_sprite's RGB888 bytes of the same rectangle straight into the caller's canvas slice (wxgui.ml:254-262, 417-424) */
CAMLprim value coh_ml_render_frame_rgb888(value ctx, value scene, value box, value out) {
  CAMLparam4(ctx, scene, box, out);
  int x = Int_val(Field(box, 0)), y = Int_val(Field(box, 1)), w = Int_val(Field(box, 2)), h = Int_val(Field(box, 3));
  if (w < 0 || h < 0) caml_failwith("Sprite.box: negative argument.");
  need(out, (int64_t)w * h * 3, "render_frame_rgb888: canvas slice too small");
  check(CTX(ctx), coh_render_frame(CTX(ctx), SCENE(scene), x, y, w, h, 0));
  check(CTX(ctx), coh_fb_read_rgb888(CTX(ctx), x, y, w, h, (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_render_frame_shape(value ctx, value scene, value update, value flags) {   /* render_frame over any update shape */
  check(CTX(ctx), coh_render_frame_shape(CTX(ctx), SCENE(scene), SHAPE(update), Int_val(flags)));
  return Val_unit;
}
CAMLprim value coh_ml_render_uncovered(value ctx) {
  CAMLparam1(ctx);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_render_uncovered(CTX(ctx), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}
/* the Sprite.sprite render_frame returns: RGBA8 per pixel of `update` in span order, in a fresh int32 Bigarray */
CAMLprim value coh_ml_fb_read_sprite(value ctx, value update) {
  CAMLparam2(ctx, update);
  CAMLlocal1(ba);
  int64_t card = 0, got = 0;
  check(CTX(ctx), coh_shape_card(CTX(ctx), SHAPE(update), &card));
  intnat dim = (intnat)card;
  ba = caml_ba_alloc(CAML_BA_INT32 | CAML_BA_C_LAYOUT, 1, NULL, &dim);
  check(CTX(ctx), coh_fb_read_sprite(CTX(ctx), SHAPE(update), (uint32_t*)Caml_ba_data_val(ba), card, &got));
  CAMLreturn(ba);
}
CAMLprim value coh_ml_read_rgba(value ctx, value box, value out) {
  CAMLparam3(ctx, box, out);
  need(out, (int64_t)Int_val(Field(box, 2)) * Int_val(Field(box, 3)) * 4, "read_rgba: buffer too small");
  check(CTX(ctx), coh_fb_read_rgba(CTX(ctx), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
/* asynchronous variant: `out` must stay alive (and is not moved: Bigarray data lives outside the OCaml heap) until read_wait */
CAMLprim value coh_ml_read_rgba_async(value ctx, value box, value out) {
  CAMLparam3(ctx, box, out);
  need(out, (int64_t)Int_val(Field(box, 2)) * Int_val(Field(box, 3)) * 4, "read_rgba_async: buffer too small");
  check(CTX(ctx), coh_fb_read_rgba_async(CTX(ctx), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_read_wait(value ctx) { check(CTX(ctx), coh_fb_read_wait(CTX(ctx))); return Val_unit; }
CAMLprim value coh_ml_read_rgb888(value ctx, value box, value out) {   /* Wxgui.plot_sprite's bytes for a rectangle */
  CAMLparam3(ctx, box, out);
  need(out, (int64_t)Int_val(Field(box, 2)) * Int_val(Field(box, 3)) * 3, "read_rgb888: canvas slice too small");
  check(CTX(ctx), coh_fb_read_rgb888(CTX(ctx), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}

/* ---- coherence across frames (engine.ml:441-493, render.ml:259-271, 1376-1438) ---- */
CAMLprim value coh_ml_scene_translate_object(value ctx, value scene, value index, value dx, value dy) {   /* Render.translate_renderobject */
  check(CTX(ctx), coh_scene_translate_object(CTX(ctx), SCENE(scene), Int_val(index), Int_val(dx), Int_val(dy)));
  return Val_unit;
}
/* Render.translate_renderobject + dirty_region + render_frame over the dirty region, as one device-side step;
 * returns the dirty pixel box (x0, y0, x1, y1) the front end re-reads with coh_ml_read_rgb888 */
CAMLprim value coh_ml_scene_drag_object(value ctx, value scene, value index, value dx, value dy) {
  CAMLparam5(ctx, scene, index, dx, dy);
  int32_t bb[4];
  check(CTX(ctx), coh_scene_drag_object(CTX(ctx), SCENE(scene), Int_val(index), Int_val(dx), Int_val(dy), 0, bb));
  CAMLreturn(tuple_of_ints(bb, 4));
}
CAMLprim value coh_ml_scene_object_shape(value ctx, value scene, value index) {   /* Render.shape_of_basicshape */
  CAMLparam3(ctx, scene, index);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_scene_object_shape(CTX(ctx), SCENE(scene), Int_val(index), &s, &m));
  CAMLreturn(pair_of_handles(s, m));
}
CAMLprim value coh_ml_dirty_filter(value ctx, value scene, value lmo, value dirty) {   /* Render.dirty_filter */
  CAMLparam4(ctx, scene, lmo, dirty);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_dirty_filter(CTX(ctx), SCENE(scene), Int_val(lmo), SHAPE(dirty), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}

/* ---- several GPUs ---- */
#define MULTI(v) ((coh_multi*)Nativeint_val(v))
static void mcheck(coh_multi* m, int rc) { if (rc) caml_failwith(coh_multi_last_error(m)); }
CAMLprim value coh_ml_multi_init(value devices) {   /* int array of device ids */
  CAMLparam1(devices);
  int32_t ids[8];
  const int n = (int)Wosize_val(devices);
  if (n < 1 || n > 8) caml_invalid_argument("coh_multi_init: 1 .. 8 devices");
  for (int k = 0; k < n; k++) ids[k] = Int_val(Field(devices, k));
  coh_multi* m = NULL;
  if (coh_multi_init(n, ids, &m)) caml_failwith(coh_multi_last_error(NULL));
  CAMLreturn(caml_copy_nativeint((intnat)m));
}
CAMLprim value coh_ml_multi_shutdown(value m) { coh_multi_shutdown(MULTI(m)); return Val_unit; }
CAMLprim value coh_ml_multi_device_count(value m) { return Val_int(coh_multi_device_count(MULTI(m))); }
CAMLprim value coh_ml_multi_ctx(value m, value i) { return caml_copy_nativeint((intnat)coh_multi_ctx(MULTI(m), Int_val(i))); }
CAMLprim value coh_ml_multi_configure(value m, value w, value h) { mcheck(MULTI(m), coh_multi_configure(MULTI(m), Int_val(w), Int_val(h), NULL)); return Val_unit; }
CAMLprim value coh_ml_multi_scene_create(value m, value objs, value n_background, value edges, value points) {
  CAMLparam5(m, objs, n_background, edges, points);
  coh_scene_t s = 0;
  mcheck(MULTI(m), coh_multi_scene_create(MULTI(m), (const coh_object*)Caml_ba_data_val(objs), (int32_t)(BA_LEN(objs) / (int64_t)sizeof(coh_object)), Int_val(n_background),
         (const int32_t*)Caml_ba_data_val(edges), (int32_t)(BA_LEN(edges) / 4), (const int32_t*)Caml_ba_data_val(points), (int32_t)(BA_LEN(points) / 2), &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_multi_scene_free(value m, value s) { mcheck(MULTI(m), coh_multi_scene_free(MULTI(m), SCENE(s))); return Val_unit; }
CAMLprim value coh_ml_multi_scene_translate_object(value m, value s, value index, value dx, value dy) {
  mcheck(MULTI(m), coh_multi_scene_translate_object(MULTI(m), SCENE(s), Int_val(index), Int_val(dx), Int_val(dy)));
  return Val_unit;
}
CAMLprim value coh_ml_multi_render_frame(value m, value s, value box, value flags) {
  CAMLparam4(m, s, box, flags);
  mcheck(MULTI(m), coh_multi_render_frame(MULTI(m), SCENE(s), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), Int_val(flags)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_multi_sync(value m) { mcheck(MULTI(m), coh_multi_sync(MULTI(m))); return Val_unit; }
CAMLprim value coh_ml_multi_read_rgba(value m, value box, value out) {
  CAMLparam3(m, box, out);
  need(out, (int64_t)Int_val(Field(box, 2)) * Int_val(Field(box, 3)) * 4, "multi_read_rgba: buffer too small");
  mcheck(MULTI(m), coh_multi_fb_read_rgba(MULTI(m), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_multi_read_rgb888(value m, value box, value out) {
  CAMLparam3(m, box, out);
  need(out, (int64_t)Int_val(Field(box, 2)) * Int_val(Field(box, 3)) * 3, "multi_read_rgb888: canvas slice too small");
  mcheck(MULTI(m), coh_multi_fb_read_rgb888(MULTI(m), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
/* one process per GPU: the 64-byte CUDA IPC handle of a framebuffer allocated for export, as a Bigarray of bytes */
CAMLprim value coh_ml_fb_alloc_shared(value ctx, value handle) {
  CAMLparam2(ctx, handle);
  need(handle, 64, "fb_alloc_shared: the handle buffer holds 64 bytes");
  check(CTX(ctx), coh_fb_alloc_shared(CTX(ctx), (uint8_t*)Caml_ba_data_val(handle)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_fb_open_peer(value ctx, value handle) {
  CAMLparam2(ctx, handle);
  void* p = NULL;
  need(handle, 64, "fb_open_peer: the handle buffer holds 64 bytes");
  check(CTX(ctx), coh_fb_open_peer(CTX(ctx), (const uint8_t*)Caml_ba_data_val(handle), &p));
  CAMLreturn(caml_copy_nativeint((intnat)p));
}

/* frame signals between processes: targets : nativeint array (pointers from fb_open_peer / fb_device_ptr) */
CAMLprim value coh_ml_frame_signal(value ctx, value targets, value slot, value epoch) {
  CAMLparam4(ctx, targets, slot, epoch);
  void* t[16];
  const long n = (long)Wosize_val(targets);
  if (n > 16) caml_invalid_argument("frame_signal: too many targets");
  for (long i = 0; i < n; i++) t[i] = (void*)Nativeint_val(Field(targets, i));
  check(CTX(ctx), coh_frame_signal(CTX(ctx), (int32_t)n, t, Int_val(slot), Int_val(epoch)));
  CAMLreturn(Val_unit);
}
CAMLprim value coh_ml_frame_wait(value ctx, value slots, value epoch) {   /* slots : int32 Array1 */
  CAMLparam3(ctx, slots, epoch);
  check(CTX(ctx), coh_frame_wait(CTX(ctx), (int32_t)BA_LEN(slots), (const int32_t*)Caml_ba_data_val(slots), Int_val(epoch)));
  CAMLreturn(Val_unit);
}

/* ---- host-side geometry (Polygon.edgelist_of_path / Brush.points_of_brushstroke for one subpath) ---- */
CAMLprim value coh_ml_host_edgelist_of_subpath(value segs, value out) {   /* segs : float64 Array1 of 9*n; returns the edge count */
  int64_t n = coh_host_edgelist_of_subpath((const double*)Caml_ba_data_val(segs), (int32_t)(BA_LEN(segs) / 9), (int32_t*)Caml_ba_data_val(out), BA_LEN(out) / 4);
  return Val_long((long)n);
}
CAMLprim value coh_ml_host_smear_points(value segs, value out) {   /* all the segments of a path; returns the point count */
  int64_t n = coh_host_smear_points((const double*)Caml_ba_data_val(segs), (int32_t)(BA_LEN(segs) / 9), (int32_t*)Caml_ba_data_val(out), BA_LEN(out) / 2);
  return Val_long((long)n);
}
CAMLprim value coh_ml_host_brush_points(value segs, value radius, value out) {
  int64_t n = coh_host_brush_points((const double*)Caml_ba_data_val(segs), (int32_t)(BA_LEN(segs) / 9), Double_val(radius), (int32_t*)Caml_ba_data_val(out), BA_LEN(out) / 2);
  return Val_long((long)n);
}

/* ---- N4 / N1: the front end's socket format (camlpy.mli) and Wxgui.refresh_window's message ----
 * kinds : int32 Array1; values, offsets : int64 Array1 (pre-order tokens, include/coherence_b200.h); strings : string */
CAMLprim value coh_ml_host_wire_marshal(value kinds, value values, value offsets, value strings) {   /* Camlpy.marshall */
  CAMLparam4(kinds, values, offsets, strings);
  CAMLlocal1(s);
  const int32_t n = (int32_t)BA_LEN(kinds);
  if (BA_LEN(values) < n || BA_LEN(offsets) < n) caml_invalid_argument("wire_marshal: token arrays of different lengths");
  const int64_t size = coh_host_wire_marshal((const int32_t*)Caml_ba_data_val(kinds), (const int64_t*)Caml_ba_data_val(values), (const int64_t*)Caml_ba_data_val(offsets), n,
                                             (const uint8_t*)String_val(strings), NULL, 0);
  if (size < 0) caml_invalid_argument("wire_marshal: not one marshallable");
  s = caml_alloc_string((size_t)size);   /* (allocation: String_val (strings) is taken again below) */
  coh_host_wire_marshal((const int32_t*)Caml_ba_data_val(kinds), (const int64_t*)Caml_ba_data_val(values), (const int64_t*)Caml_ba_data_val(offsets), n,
                        (const uint8_t*)String_val(strings), (uint8_t*)Bytes_val(s), size);
  CAMLreturn(s);
}
/* Camlpy.unmarshall: (bytes taken, tokens); taken = 0: the message is not complete yet; Failure "Invalid_data" */
CAMLprim value coh_ml_host_wire_unmarshal(value str, value kinds, value values, value offsets) {
  CAMLparam4(str, kinds, values, offsets);
  int32_t nt = 0, cap = (int32_t)BA_LEN(kinds);
  int64_t taken = 0;
  if (BA_LEN(values) < cap || BA_LEN(offsets) < cap) caml_invalid_argument("wire_unmarshal: token arrays of different lengths");
  if (coh_host_wire_unmarshal((const uint8_t*)String_val(str), (int64_t)caml_string_length(str), (int32_t*)Caml_ba_data_val(kinds), (int64_t*)Caml_ba_data_val(values),
                              (int64_t*)Caml_ba_data_val(offsets), cap, &nt, &taken) != 0) caml_failwith("Invalid_data");
  int32_t r[2] = {(int32_t)taken, nt};
  CAMLreturn(tuple_of_ints(r, 2));
}
/* the bytes in front of the pixels of a RefreshWindow message and the size of the whole message; ("", 0): nothing to send */
CAMLprim value coh_ml_host_wire_refresh_window(value window, value rect) {   /* rect = (xmin, ymin, xmax, ymax), inclusive */
  CAMLparam2(window, rect);
  CAMLlocal2(h, r);
  uint8_t hdr[64]; int32_t hl = 0;
  const int64_t total = coh_host_wire_refresh_window(Int_val(window), Int_val(Field(rect, 0)), Int_val(Field(rect, 1)), Int_val(Field(rect, 2)), Int_val(Field(rect, 3)), hdr, &hl);
  if (total < 0) caml_invalid_argument("refresh_window: not a rectangle");
  h = caml_alloc_string((size_t)hl);
  memcpy(Bytes_val(h), hdr, (size_t)hl);
  r = caml_alloc_tuple(2);
  Store_field(r, 0, h);
  Store_field(r, 1, Val_long((long)total));
  CAMLreturn(r);
}
/* Wxgui.refresh_window's whole message for a dirty rectangle, pixels from the GPU framebuffer ("" = nothing to send) */
CAMLprim value coh_ml_wire_refresh_window(value ctx, value window, value rect) {
  CAMLparam3(ctx, window, rect);
  CAMLlocal1(s);
  const int32_t w = Int_val(window), x0 = Int_val(Field(rect, 0)), y0 = Int_val(Field(rect, 1)), x1 = Int_val(Field(rect, 2)), y1 = Int_val(Field(rect, 3));
  int64_t len = 0;
  check(CTX(ctx), coh_wire_refresh_window(CTX(ctx), w, x0, y0, x1, y1, NULL, 0, &len));
  s = caml_alloc_string((size_t)len);
  if (len > 0) check(CTX(ctx), coh_wire_refresh_window(CTX(ctx), w, x0, y0, x1, y1, (uint8_t*)Bytes_val(s), len, &len));   /* (blocks; the runtime lock is held: s does not move) */
  CAMLreturn(s);
}
