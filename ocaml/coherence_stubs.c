/* coherence_stubs.c — OCaml externals over the C ABI (include/coherence_b200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no OCaml toolchain here (no ocamlfind,
 * no caml/mlvalues.h).  Delivered as reviewed source; build with
 *   ocamlfind ocamlopt -package bigarray -I include -c ocaml/coherence_stubs.c
 * and link with -cclib -lcoherence_b200.
 *
 * Conventions: bulk data arrives in Bigarray.Array1 (c_layout, int32 / int8_unsigned); handles
 * are boxed int64 (Int64.t); a non-zero status raises Failure (coh_last_error), which is the
 * reference's own error convention (failwith).  The stubs never retain OCaml pointers across
 * calls, so the GC may move values freely; device objects are freed by explicit calls wired to
 * Gc.finalise in coherence_gpu.ml.  The caller is the single OCaml thread; calls block until the
 * requested host bytes are ready, so the runtime lock is not released.
 */
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>
#include <string.h>
#include "coherence_b200.h"

#define CTX(v) ((coh_ctx*)Nativeint_val(v))
static void check(coh_ctx* c, int rc) { if (rc) caml_failwith(coh_last_error(c)); }

CAMLprim value coh_ml_init(value device) {
  CAMLparam1(device);
  coh_ctx* c = NULL;
  if (coh_init(Int_val(device), &c)) caml_failwith(coh_last_error(NULL));
  CAMLreturn(caml_copy_nativeint((intnat)c));
}
CAMLprim value coh_ml_shutdown(value ctx) { coh_shutdown(CTX(ctx)); return Val_unit; }

/* Colour.colour (31-bit int) <-> RGBA8 word */
CAMLprim value coh_ml_rgba8_of_colour(value c) { return caml_copy_int32((int32_t)coh_rgba8_of_colour((int32_t)Long_val(c))); }
CAMLprim value coh_ml_colour_of_rgba8(value w) { return Val_long(coh_colour_of_rgba8((uint32_t)Int32_val(w))); }

/* Polygon.shapeminshape_of_unsorted_edgelist: edges : (int32, c_layout) Array1 of 4*n */
CAMLprim value coh_ml_shapeminshape(value ctx, value edges, value winding) {
  CAMLparam3(ctx, edges, winding);
  CAMLlocal1(pair);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_shapeminshape_of_edgelist(CTX(ctx), (const int32_t*)Caml_ba_data_val(edges),
        (int32_t)(Caml_ba_array_val(edges)->dim[0] / 4), Int_val(winding), &s, &m));
  pair = caml_alloc_tuple(2);
  Store_field(pair, 0, caml_copy_int64((int64_t)s));
  Store_field(pair, 1, caml_copy_int64((int64_t)m));
  CAMLreturn(pair);
}
/* Sprite.shape export: returns the flat (y, nspans, (x,len)...) records in a fresh Bigarray */
CAMLprim value coh_ml_shape_export(value ctx, value shape) {
  CAMLparam2(ctx, shape);
  int64_t n = 0, got = 0;
  check(CTX(ctx), coh_shape_export_size(CTX(ctx), (coh_shape_t)Int64_val(shape), &n));
  intnat dim = (intnat)n;
  value ba = caml_ba_alloc(CAML_BA_INT32 | CAML_BA_C_LAYOUT, 1, NULL, &dim);
  check(CTX(ctx), coh_shape_export(CTX(ctx), (coh_shape_t)Int64_val(shape), (int32_t*)Caml_ba_data_val(ba), n, &got));
  CAMLreturn(ba);
}
CAMLprim value coh_ml_shape_import(value ctx, value flat) {
  CAMLparam2(ctx, flat);
  coh_shape_t s = 0;
  check(CTX(ctx), coh_shape_import(CTX(ctx), (const int32_t*)Caml_ba_data_val(flat), Caml_ba_array_val(flat)->dim[0], &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_shape_free(value ctx, value s) { coh_shape_free(CTX(ctx), (coh_shape_t)Int64_val(s)); return Val_unit; }
#define BINOP(name, fn)                                                              \
  CAMLprim value name(value ctx, value a, value b) {                                 \
    CAMLparam3(ctx, a, b);                                                           \
    coh_shape_t o = 0;                                                               \
    check(CTX(ctx), fn(CTX(ctx), (coh_shape_t)Int64_val(a), (coh_shape_t)Int64_val(b), &o)); \
    CAMLreturn(caml_copy_int64((int64_t)o));                                         \
  }
BINOP(coh_ml_shape_union, coh_shape_union)              /* Sprite.( ||| ) */
BINOP(coh_ml_shape_difference, coh_shape_difference)    /* Sprite.( --- ) */
BINOP(coh_ml_shape_intersection, coh_shape_intersection)/* Sprite.( &&& ) */
CAMLprim value coh_ml_shape_bloat(value ctx, value a, value m, value n) {
  CAMLparam4(ctx, a, m, n);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_shape_bloat(CTX(ctx), (coh_shape_t)Int64_val(a), Int_val(m), Int_val(n), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}

/* Render: objs is a Bigarray of bytes holding n packed coh_object records built by
 * Coherence_gpu.flatten_scene (closures -> descriptors happens on the OCaml side). */
CAMLprim value coh_ml_scene_create(value ctx, value objs, value n_background, value edges, value points) {
  CAMLparam5(ctx, objs, n_background, edges, points);
  coh_scene_t s = 0;
  check(CTX(ctx), coh_scene_create(CTX(ctx), (const coh_object*)Caml_ba_data_val(objs),
        (int32_t)(Caml_ba_array_val(objs)->dim[0] / sizeof(coh_object)), Int_val(n_background),
        (const int32_t*)Caml_ba_data_val(edges), (int32_t)(Caml_ba_array_val(edges)->dim[0] / 4),
        (const int32_t*)Caml_ba_data_val(points), (int32_t)(Caml_ba_array_val(points)->dim[0] / 2), &s));
  CAMLreturn(caml_copy_int64((int64_t)s));
}
CAMLprim value coh_ml_scene_free(value ctx, value s) { coh_scene_free(CTX(ctx), (coh_scene_t)Int64_val(s)); return Val_unit; }
CAMLprim value coh_ml_fb_configure(value ctx, value w, value h, value y0, value y1) {
  check(CTX(ctx), coh_fb_configure(CTX(ctx), Int_val(w), Int_val(h), Int_val(y0), Int_val(y1)));
  return Val_unit;
}
/* Render.render_frame over update = Sprite.box x y w h, then plot_sprite's RGB888 bytes of the
 * same rectangle straight into the caller's canvas slice (wxgui.ml:254-262, 417-424). */
CAMLprim value coh_ml_render_frame_rgb888(value ctx, value scene, value box, value out) {
  CAMLparam4(ctx, scene, box, out);
  int x = Int_val(Field(box, 0)), y = Int_val(Field(box, 1)), w = Int_val(Field(box, 2)), h = Int_val(Field(box, 3));
  check(CTX(ctx), coh_render_frame(CTX(ctx), (coh_scene_t)Int64_val(scene), x, y, w, h, 0));
  check(CTX(ctx), coh_fb_read_rgb888(CTX(ctx), x, y, w, h, (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}

/* ---- coherence across frames (engine.ml:441-493, render.ml:259-271, 1376-1438, cache.mli) ---- */
CAMLprim value coh_ml_cache_configure(value ctx, value on, value bytes) {   /* Cache.usecache / Cache.setsize */
  check(CTX(ctx), coh_cache_configure(CTX(ctx), Bool_val(on), Int64_val(bytes)));
  return Val_unit;
}
CAMLprim value coh_ml_cache_clear(value ctx) { check(CTX(ctx), coh_cache_clear(CTX(ctx))); return Val_unit; }
/* Render.translate_renderobject + dirty_region + render_frame over the dirty region, as one device-side step;
 * returns the dirty pixel box (x0, y0, x1, y1) the front end re-reads with coh_ml_read_rgb888 */
CAMLprim value coh_ml_scene_drag_object(value ctx, value scene, value index, value dx, value dy) {
  CAMLparam5(ctx, scene, index, dx, dy);
  CAMLlocal1(r);
  int32_t bb[4];
  check(CTX(ctx), coh_scene_drag_object(CTX(ctx), (coh_scene_t)Int64_val(scene), Int_val(index), Int_val(dx), Int_val(dy), 0, bb));
  r = caml_alloc_tuple(4);
  for (int k = 0; k < 4; k++) Store_field(r, k, Val_int(bb[k]));
  CAMLreturn(r);
}
CAMLprim value coh_ml_scene_object_shape(value ctx, value scene, value index) {   /* Render.shape_of_basicshape */
  CAMLparam3(ctx, scene, index);
  CAMLlocal1(r);
  coh_shape_t s = 0, m = 0;
  check(CTX(ctx), coh_scene_object_shape(CTX(ctx), (coh_scene_t)Int64_val(scene), Int_val(index), &s, &m));
  r = caml_alloc_tuple(2);
  Store_field(r, 0, caml_copy_int64((int64_t)s));
  Store_field(r, 1, caml_copy_int64((int64_t)m));
  CAMLreturn(r);
}
CAMLprim value coh_ml_dirty_filter(value ctx, value scene, value lmo, value dirty) {   /* Render.dirty_filter */
  CAMLparam4(ctx, scene, lmo, dirty);
  coh_shape_t o = 0;
  check(CTX(ctx), coh_dirty_filter(CTX(ctx), (coh_scene_t)Int64_val(scene), Int_val(lmo), (coh_shape_t)Int64_val(dirty), &o));
  CAMLreturn(caml_copy_int64((int64_t)o));
}
CAMLprim value coh_ml_render_frame_shape(value ctx, value scene, value update) {   /* render_frame over any update shape */
  check(CTX(ctx), coh_render_frame_shape(CTX(ctx), (coh_scene_t)Int64_val(scene), (coh_shape_t)Int64_val(update), 0));
  return Val_unit;
}
CAMLprim value coh_ml_read_rgb888(value ctx, value box, value out) {   /* Wxgui.plot_sprite's bytes for a rectangle */
  CAMLparam3(ctx, box, out);
  check(CTX(ctx), coh_fb_read_rgb888(CTX(ctx), Int_val(Field(box, 0)), Int_val(Field(box, 1)), Int_val(Field(box, 2)), Int_val(Field(box, 3)), (uint8_t*)Caml_ba_data_val(out)));
  CAMLreturn(Val_unit);
}
/* Convolve.convolve_sprite kernel sprite: sprite = (shape handle, RGBA8 per pixel in span order) */
CAMLprim value coh_ml_convolve_sprite(value ctx, value kind_r, value shape, value rgba_in, value rgba_out) {
  CAMLparam5(ctx, kind_r, shape, rgba_in, rgba_out);
  coh_shape_t o = 0;
  int64_t n = 0;
  check(CTX(ctx), coh_convolve_sprite(CTX(ctx), Int_val(Field(kind_r, 0)), Int_val(Field(kind_r, 1)), (coh_shape_t)Int64_val(shape),
        (const uint32_t*)Caml_ba_data_val(rgba_in), &o, (uint32_t*)Caml_ba_data_val(rgba_out), Caml_ba_array_val(rgba_out)->dim[0], &n));
  CAMLreturn(caml_copy_int64((int64_t)o));
}
