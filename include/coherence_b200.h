/* coherence_b200.h — C ABI of the B200-native Coherence raster hot path.
 *
 * This is the drop-in boundary: OCaml `external` stubs (ocaml/coherence_stubs.c, see
 * INTEGRATION.md) bind exactly these entry points behind the reference's own module
 * signatures.  There is no FFI in the reference today; each entry point names the
 * OCaml interface it replaces (file:line in /root/reference).
 *
 * Conventions
 *  - plain C, `extern "C"`, caller-owned contiguous buffers (OCaml Bigarray.Array1
 *    c_layout), opaque 64-bit handles for device-resident objects.
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    available from coh_last_error().  The OCaml stub raises `Failure msg`
 *    (reference convention: failwith, 98 sites; e.g. sprite.ml:463, render.ml:1274).
 *  - colours cross the boundary as RGBA8 words r | g<<8 | b<<16 | a<<24,
 *    premultiplied (r,g,b <= a), the decoded form of the reference's 31-bit
 *    `Colour.colour` (colour.ml:66-244); coh_colour_* convert.
 *  - geometry crosses the boundary AFTER the affine transform, as integer edges in
 *    sub-pixel bins (Coord.sub_of_float, coord.ml:47): int32 {x0, y0, x1, y1}.
 *  - span sets ("shapes", sprite.ml:46-54) are exported in canonical order as
 *    int32 records: for every non-empty row in increasing y:  y, nspans, then nspans
 *    pairs (x, len).  Rows group into the reference's vspans by consecutive y.
 *  - there is NO CPU fallback: every compute entry point fails with an error if no
 *    CUDA device is usable.
 */
#ifndef COHERENCE_B200_H
#define COHERENCE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct coh_ctx coh_ctx;
typedef uint64_t coh_shape_t; /* device-resident span set; 0 == Sprite.NullShape */
typedef uint64_t coh_scene_t; /* device-resident scene */

/* ---- scene description (render.ml:19-75 `renderobject`, flattened depth-first) ---- */
enum {
  COH_OBJ_PATH = 0,        /* Basic (fill, Path p), edges = Polygon.edgelist_of_path      */
  COH_OBJ_PRIMITIVE = 1,   /* Primitive (colour, HLine|VLine|Rectangle), render.ml:556-586 */
  COH_OBJ_GROUP_BEGIN = 2, /* Group scene ... */
  COH_OBJ_GROUP_END = 3,   /* ... end of the innermost open group                          */
  COH_OBJ_BRUSH = 4,       /* Basic (fill, Brushstroke ((opacity, Gaussian r), path))      */
  COH_OBJ_CPG = 5,         /* Basic (fill, CPG (op, Path a, Path b)), render.ml:17-18, 522-528, 867-981 */
  COH_OBJ_FILTER = 6       /* Filter {geometry = Basic (fill, Path); reading_scene; filter} (render.ml:37-48, 1080-1131) */
};
/* Filters (filters.ml).  The reference passes closures; across the ABI a filter is a descriptor:
 *   HOLE        filters.ml:216-224  reading scene = [], filter = identity
 *   MONOCHROME  filters.ml:229-238  reading scene = the objects below, filter = sprite_map Colour.monochrome
 *   BLUR        filters.ml:243-258  reading scene = the objects below read in bloat (2r+1) (2r+1) shape,
 *                                   filter = Convolve.convolve_sprite_in_shape kernel (filter_kernel = COH_CONV_* | r << 8)
 *   SCENE       affine / rgb / wireframe / swapdepth (filters.ml:105-212, 271-285, 305-332): the caller builds the
 *               modified scene (a host-side rewrite of the objects below) and passes it as a reading-scene
 *               group; filter = identity.  first2 = index (in the objs array) of that group's GROUP_BEGIN.
 *   MINUS       filters.ml:289-303  a single-object hole: the object that follows the filter in its list is cut away
 *                                   inside the filter — reading scene = the objects below but that first one, read (and the
 *                                   filter applied) only in shape (filter) ∩ shape (first object below); filter = identity
 *   SMEAR       filters.ml:201-217  smear along a brush stroke (Brush.smear, brush.ml:235-331): the geometry is the
 *                                   stroke's dummy (first / count = its stamp points in the points array, brush_radius,
 *                                   brush_opacity = the brush), the scene below is read in bloat r r shape, and
 *                                   first2 / count2 = the integer points of Brush.find_smear_directions, consecutive
 *                                   duplicates dropped (coh_host_smear_points), in the points array
 * A reading-scene group is a top-level GROUP_BEGIN of the scene section with filter_kind =
 * COH_FILTER_READING_SCENE; such groups come after every ordinary scene object and are never drawn
 * on their own.  A filter object may carry an alias offset (dx, dy) and may be a member of a group: such a group's members
 * are rendered once, as a scene of their own, into the group's canvas (like Convolved (kernel, Group members)), so they
 * take plain fills only; SCENE filters must be top-level members of the scene list.  The geometry of a filter is the path
 * with a plain fill described by the filter's own record (first / count / winding / colour0), or — cpg_op = COH_GEOM_NEXT —
 * the OBJECT THAT FOLLOWS the filter in the list (one record, or a whole GROUP_BEGIN .. GROUP_END): a brush stroke
 * (examples.ml:279-300), a Convolved path (engine.ml:33-36: soft-edged lenses), a CPG, a stroked path, a group; plain
 * fills.  That object is consumed by the filter: it is not drawn, and the filter's "objects below" begin after it.  Its
 * shape is the filter's shape and the alpha of its sprite the filter's matte (render.ml:1099). */
enum { COH_FILTER_NONE = 0, COH_FILTER_HOLE = 1, COH_FILTER_MONOCHROME = 2, COH_FILTER_BLUR = 3, COH_FILTER_SCENE = 4, COH_FILTER_MINUS = 5, COH_FILTER_SMEAR = 6,
       COH_FILTER_READING_SCENE = 100 };
enum { COH_CPG_UNION = 0, COH_CPG_INTERSECTION = 1, COH_CPG_SUBTRACTION = 2, COH_CPG_EXCLUSIVEOR = 3 };
enum { COH_NONZERO = 0, COH_EVENODD = 1 };              /* Pdfgraphics.winding_rule */
enum { COH_FILL_PLAIN = 0, COH_FILL_AXIAL = 1, COH_FILL_RADIAL = 2 }; /* fill.ml:62,77,112 */
enum { COH_FILL_EXT_S = 1, COH_FILL_EXT_E = 2 };
enum { COH_GEOM_PATH = 0, COH_GEOM_NEXT = 1 };  /* FILTER objects carry it in `cpg_op`: the geometry is the filter record's own path, or the object that follows */
enum { COH_BRUSH_GAUSSIAN = 0, COH_BRUSH_DUMMY = 1 };  /* Brush.brushkind (brush.ml:14-16); BRUSH objects carry it in `winding` */
enum { COH_CONV_UNIT = 1, COH_CONV_GAUSSIAN = 2 };  /* convolve.ml:19-22,37-70 (UnitKernel r / XYKernel from mkgaussian r) */

typedef struct coh_object {
  int32_t kind;        /* COH_OBJ_* */
  int32_t winding;     /* COH_NONZERO / COH_EVENODD (PATH); BRUSH: COH_BRUSH_GAUSSIAN, or COH_BRUSH_DUMMY = Dummy (r, r) with
                          r = brush_radius (an integer): the whole shape of the stroke in white, whatever the fill
                          (brush.ml:178-181; Brush.mkdummy) */
  int32_t first;       /* PATH: first edge; BRUSH: first point */
  int32_t count;       /* PATH: number of edges; BRUSH: number of points */
  int32_t fill_kind;   /* COH_FILL_* */
  uint32_t colour0;    /* plain colour / gradient start `cs`; PRIMITIVE colour (RGBA8 premultiplied) */
  uint32_t colour1;    /* gradient end `ce` */
  int32_t fill_flags;  /* COH_FILL_EXT_S | COH_FILL_EXT_E */
  int32_t pretrans;    /* -1: Over; 0..255: PreTrans (v, Over) with toint (v *. 255.) (render.ml:1295-1298) */
  int32_t dx, dy;      /* integer pixel translation alias (Cache.addtranslation, cache.ml:423-436) */
  int32_t bounds[4];   /* bounds_of_basicshape xmin,xmax,ymin,ymax (render.ml:377-437); used for the
                          trivial reject of render.ml:1270-1279 only where it is row-local */
  int32_t prim[4];     /* PRIMITIVE: inclusive pixel box x0,y0,x1,y1 (toint of the float rectangle) */
  int32_t prim_null;   /* PRIMITIVE: 1 for a zero-length HLine/VLine (NullShape) */
  int32_t convolve;    /* PATH: 0, or Convolved (kernel, Basic (fill, Path p)) (render.ml:63, 1023-1052):
                          COH_CONV_UNIT | r << 8  = Convolve.mkunit r,  COH_CONV_GAUSSIAN | r << 8 = Convolve.mkgaussian r.
                          GROUP_BEGIN: 0, or Convolved (kernel, Group members) — the members up to the matching GROUP_END
                          (plain fills only; no filters) are rendered once into the object's canvas and convolved there */
  int32_t sprite_winding; /* PATH: 0 = the AA sprite uses `winding`; 1 + rule otherwise.  Basic (fill, StrokedPath (p, spec)),
                          with edges = Shapes.strokepath spec p, takes its shape with NonZero (render.ml:510) but
                          its sprite with EvenOdd (render.ml:1018): winding = COH_NONZERO, sprite_winding = 1 + COH_EVENODD */
  int64_t id;          /* cache key (Id.idset); < 0: fresh id each render, never cached */
  double fparam[6];    /* AXIAL: x0,y0,x1,y1;  RADIAL: cx,cy, px,py, p'x,p'y (fill.ml:77,112) */
  double brush_opacity; /* BRUSH: opacity in 0..1 */
  double brush_radius;  /* BRUSH: Gaussian radius */
  int32_t first2;       /* CPG: edges of operand b (operand a uses first / count / winding) */
  int32_t count2;
  int32_t winding2;     /* CPG: winding rule of operand b */
  int32_t cpg_op;       /* CPG: COH_CPG_* */
  int32_t filter_kind;  /* FILTER: COH_FILTER_*; GROUP_BEGIN: COH_FILTER_READING_SCENE or 0 */
  int32_t filter_kernel;/* FILTER BLUR: COH_CONV_UNIT | COH_CONV_GAUSSIAN, radius << 8 */
} coh_object;

/* ---- lifecycle ---- */
/* One context owns one GPU and one horizontal band of scanlines [band_y0, band_y1).
 * device < 0 selects the current device.  Fails when no CUDA device is present. */
int coh_init(int device, coh_ctx** out);
int coh_shutdown(coh_ctx* ctx);
const char* coh_last_error(coh_ctx* ctx); /* ctx may be NULL for init errors */
int coh_device_name(coh_ctx* ctx, char* buf, int cap);
/* the CUDA stream every kernel of this context is launched on (for event timing) */
void* coh_stream(coh_ctx* ctx);
/* run on a caller-owned CUDA stream instead (e.g. the one a collective library uses) */
int coh_set_stream(coh_ctx* ctx, void* cuda_stream);
/* number of kernel launches issued by this context since creation */
int64_t coh_launch_count(coh_ctx* ctx);
/* per-kernel timing with CUDA events on the launching stream: average duration of the fused
 * walker kernel and of the binning kernels over the frames rendered since it was switched on */
int coh_set_timing(coh_ctx* ctx, int32_t on);
int coh_get_timing(coh_ctx* ctx, double* walk_ms_avg, double* bin_ms_avg, int64_t* frames);

/* Tuning and test options (none changes a result): "walk_h" 0 | 1 | 4 | 16 — rows per walker work item (0: chosen per
 * pass); "fused" -1 | 0 | 1 — force the three-phase frame (0) or the fused walker (1); "aa_general" 0 | 1 — every
 * antialiased pair through the general bit-row kernel instead of the interval form; "bin_cache" 0 | 1 — keep the
 * whole-frame cell binning with the scene; "comp_rows" 0 | 1 — flat scenes composite three-phase frames with the row
 * compositor instead of the walker; "fork_prefill" 0 | 1 — the background prefill of a three-phase frame runs on a
 * second stream beside the scan kernels.  coh_init reads COH_WALK_H, COH_FUSED, COH_AA_GENERAL from the environment once. */
int coh_set_option(coh_ctx* ctx, const char* name, int32_t value);

/* ---- colour codec (colour.ml:99-172 colour_of_rgba / rgba_of_colour) ---- */
int32_t coh_colour_of_rgba8(uint32_t rgba8);
uint32_t coh_rgba8_of_colour(int32_t colour);

/* ---- Polygon (polygon.mli:44-59) ---- */
/* Polygon.shapeminshape_of_unsorted_edgelist edges winding  (polygon.ml:608-609):
 * scan-converts one edge list into (shape, minshape), both device resident. */
int coh_shapeminshape_of_edgelist(coh_ctx* ctx, const int32_t* edges, int32_t n_edges, int32_t winding,
                                  coh_shape_t* shape, coh_shape_t* minshape);
/* Antialiased coverage of Polygon.polygon_sprite_edgelist (polygon.ml:694-746) without
 * the fill: opacity 0..255 for every pixel of `shp`, in canonical span order. */
int coh_polygon_opacity(coh_ctx* ctx, const int32_t* edges, int32_t n_edges, int32_t winding,
                        coh_shape_t shp, uint8_t* opacity_out, int64_t cap, int64_t* n_out);
/* Polygon.polygon_sprite_edgelist fill shp edges winding (polygon.ml:729-746): RGBA8 per
 * pixel of `shp` in canonical span order.  `fill` uses the fill_* / colour* / fparam fields. */
int coh_polygon_sprite(coh_ctx* ctx, const coh_object* fill, const int32_t* edges, int32_t n_edges,
                       int32_t winding, coh_shape_t shp, uint32_t* rgba_out, int64_t cap, int64_t* n_out);

/* ---- Sprite span-set algebra (sprite.mli:83-136,173-175) ---- */
int coh_shape_box(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, coh_shape_t* out); /* Sprite.box, sprite.ml:462 */
int coh_shape_import(coh_ctx* ctx, const int32_t* flat, int64_t n, coh_shape_t* out);
int coh_shape_export_size(coh_ctx* ctx, coh_shape_t s, int64_t* n_int32);
int coh_shape_export(coh_ctx* ctx, coh_shape_t s, int32_t* flat, int64_t cap, int64_t* n_out);
int coh_shape_bounds(coh_ctx* ctx, coh_shape_t s, int32_t box[4], int32_t* is_null); /* x0,y0,x1,y1 (boxshape, sprite.ml:542) */
int coh_shape_card(coh_ctx* ctx, coh_shape_t s, int64_t* npixels);
int coh_shape_free(coh_ctx* ctx, coh_shape_t s);
int coh_shape_union(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out);        /* `|||` sprite.ml:1275 */
int coh_shape_difference(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out);   /* `---` sprite.ml:1483 */
int coh_shape_intersection(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, coh_shape_t* out); /* `&&&` sprite.ml:1623 */
int coh_shape_translate(coh_ctx* ctx, coh_shape_t a, int32_t dx, int32_t dy, coh_shape_t* out); /* sprite.ml:476 */
int coh_shape_bloat(coh_ctx* ctx, coh_shape_t a, int32_t m, int32_t n, coh_shape_t* out); /* sprite.ml:1857 */
int coh_shape_erode(coh_ctx* ctx, coh_shape_t a, int32_t m, int32_t n, coh_shape_t* out); /* sprite.ml:1867 */

/* ---- Sprite operations on whole sprites (sprite.mli:96-125) ----
 * A sprite crosses the boundary as its shape (a device span set) plus one RGBA8 word per pixel in canonical span
 * order.  Sprite.translate_sprite (sprite.mli:106) is coh_shape_translate on the shape: the pixels do not change. */
int coh_shape_intersects(coh_ctx* ctx, coh_shape_t a, coh_shape_t b, int32_t* yes);   /* Sprite.shape_intersects, sprite.ml:1661 */
/* Sprite.portion spr shp (sprite.ml:642-721): fails ("portion_spanline: bad input") unless shp lies inside the sprite's shape */
int coh_sprite_portion(coh_ctx* ctx, coh_shape_t shape, const uint32_t* rgba, coh_shape_t sub, uint32_t* rgba_out, int64_t cap, int64_t* n_out);
/* Sprite.fillshape shp fill (sprite.ml:158-175); `fill` uses the fill_* / colour* / fparam fields */
int coh_sprite_fillshape(coh_ctx* ctx, coh_shape_t shape, const coh_object* fill, uint32_t* rgba_out, int64_t cap, int64_t* n_out);
/* Sprite.sprite_map f spr (sprite.ml:358-374); the closure becomes an enumerated colour function (colour.ml:266-304) */
enum { COH_MAP_MONOCHROME = 0, COH_MAP_DISSOLVE = 1 /* ~delta:arg */, COH_MAP_RED_CHANNEL = 2, COH_MAP_GREEN_CHANNEL = 3, COH_MAP_BLUE_CHANNEL = 4 };
int coh_sprite_map(coh_ctx* ctx, int32_t op, int32_t arg, const uint32_t* rgba_in, int64_t n, uint32_t* rgba_out);
/* Sprite.map_coords (fun x y c -> dissolve (fill x y) ~delta:(alpha c)) spr (sprite.ml:307-356; render.ml:976-981) */
int coh_sprite_map_coords_fill(coh_ctx* ctx, coh_shape_t shape, const coh_object* fill, const uint32_t* rgba_in, uint32_t* rgba_out, int64_t cap, int64_t* n_out);

/* ---- Convolve (convolve.mli:28-40) ----
 * Convolve.convolve_sprite kernel sprite (convolve.ml:239-258) with kernel = mkunit r / mkgaussian r.  A sprite
 * crosses the boundary as its shape plus one RGBA8 word per pixel in canonical span order; the result
 * lives on bloat r r (shape) (returned in *out_shape) and is written the same way. */
int coh_convolve_sprite(coh_ctx* ctx, int32_t kernel_kind, int32_t r, coh_shape_t shape, const uint32_t* rgba_in,
                        coh_shape_t* out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out);

/* ---- Cache (cache.mli:32-48): span sets resident in HBM, keyed by Id.idset ---- */
int coh_cache_configure(coh_ctx* ctx, int32_t usecache, int64_t max_bytes); /* Cache.usecache / setsize (default 50 MiB, cache.ml:73) */
int coh_cache_clear(coh_ctx* ctx);                                           /* Cache.clear */
int coh_cache_stats(coh_ctx* ctx, int64_t out[4]);                           /* shape hits, misses, bytes, entries (cache.ml:24-38) */
/* Partial sprites (Cache.addsprite / getsprite, cache.ml:328-367, 390-407; render.ml:1169-1242) are kept per scene: a
 * top-level Group with an id >= 0 (plain-filled paths and primitives inside) owns an RGBA8 canvas and a pshape plane in
 * HBM; a frame renders only what of it is not cached yet, and a drag reads the cached sprite translated.
 * out = frames served from the sprites alone, frames that rendered into them, bytes resident, cached objects. */
int coh_cache_sprite_stats(coh_ctx* ctx, coh_scene_t scene, int64_t out[4]);
int coh_cache_addshape(coh_ctx* ctx, int64_t id, coh_shape_t shape, coh_shape_t minshape); /* Cache.addshape: copies kept, cache.ml:280 */
int coh_cache_getshape(coh_ctx* ctx, int64_t id, coh_shape_t* shape, coh_shape_t* minshape, int32_t* found); /* Cache.getshape, cache.ml:370 */
int coh_cache_addtranslation(coh_ctx* ctx, int64_t id, int64_t target, int32_t dx, int32_t dy); /* cache.ml:423 */
/* Render.plaindirty (plain != 0) / alldirty (render.ml:1376-1391) */
int coh_dirty_region(coh_ctx* ctx, coh_shape_t shp_o, coh_shape_t minshp_o, coh_shape_t shp_n, coh_shape_t minshp_n,
                     coh_shape_t u, int32_t plain, coh_shape_t* out);

/* ---- Render (render.mli:211-217) ---- */
/* Upload a scene (flattened object list, head = front-most) with its edge and brush
 * point pools.  Edges: int32[n_edges][4] = x0,y0,x1,y1 sub-bins.  Points: int32[n][2].
 * The LAST n_background objects are render_frame's (view.pages @ view.background) list
 * (render.ml:1364), the others its scene list (render.ml:1357-1363). */
int coh_scene_create(coh_ctx* ctx, const coh_object* objs, int32_t n_objs, int32_t n_background,
                     const int32_t* edges, int32_t n_edges, const int32_t* points, int32_t n_points,
                     coh_scene_t* out);
int coh_scene_free(coh_ctx* ctx, coh_scene_t s);
/* Size of the device framebuffer (RGBA8, W x H, pixel (0,0) first) and the scanline band
 * [band_y0, band_y1) this context renders; rows outside the band are left untouched. */
int coh_fb_configure(coh_ctx* ctx, int32_t width, int32_t height, int32_t band_y0, int32_t band_y1);
/* Render into a caller-owned device buffer of W*H RGBA8 words instead (band gather by NCCL
 * happens on buffers the collective library knows); NULL goes back to a buffer owned by the context. */
int coh_fb_attach(coh_ctx* ctx, void* device_rgba8);
/* Render.render_frame lmo view update (render.ml:1345-1365) with update = Sprite.box ux uy uw uh:
 * scene pass over (pages @ background) pass, composited front to back with hidden-surface
 * set subtraction, into the device framebuffer.  Pixels of the update box not reached by
 * any object become clear (0).  Asynchronous on coh_stream(); Render.render_simple_scene
 * (render.ml:1368-1370) is the same call on a scene created with n_background = 0. */
enum { COH_RENDER_RECORD_U = 1 };
int coh_render_frame(coh_ctx* ctx, coh_scene_t scene, int32_t ux, int32_t uy, int32_t uw, int32_t uh,
                     int32_t flags);
/* Render.dirty_filter lmo initial_dirty scene (render.ml:1418-1438): the dirty functions of the filters in
 * front of the last-moved object (lmo_index, an index into the objs array; -1 = every filter), composed from the
 * one nearest the object to the front-most: hole / monochrome / caller-built reading scenes use nulldirty
 * (filters.ml:9-12; the shim composes the transform-based dirty functions of affine / rgb / wireframe itself with
 * the coh_shape_* operations), blur uses bloatdirty r r (filters.ml:63-75). */
int coh_dirty_filter(coh_ctx* ctx, coh_scene_t scene, int32_t lmo_index, coh_shape_t initial_dirty, coh_shape_t* out);
/* One step of an interactive drag (engine.ml:441-493 around render.ml:259-271, 1376-1400 and 1345-1365):
 * the object (or group) becomes an alias of its cached self moved by (dx, dy) whole pixels, the dirty region
 * dirty_region obj obj' = plaindirty | alldirty is formed from the cached, HBM-resident span sets, intersected
 * with the framebuffer rectangle, and render_frame runs over exactly that region.  Nothing is synchronised or
 * copied to the host; dirty_bbox (may be NULL) receives the pixel box x0, y0, x1, y1 (inclusive, clipped to the
 * frame; x1 < x0 when nothing is dirty) that a front end would re-read with coh_fb_read_rgb888.
 * Equivalent to coh_scene_object_shape + coh_scene_translate_object + coh_scene_object_shape +
 * coh_dirty_region + coh_render_frame_shape, without materialising the region as a span set. */
int coh_scene_drag_object(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy, int32_t flags,
                          int32_t dirty_bbox[4]);
/* Render.render_frame over an arbitrary update shape — the dirty region that engine.ml:224-252
 * (force_update) passes after a change; pixels outside the shape keep their previous value. */
int coh_render_frame_shape(coh_ctx* ctx, coh_scene_t scene, coh_shape_t update, int32_t flags);
/* Render.translate_renderobject dx dy obj (render.ml:259-271) on the obj_index-th object of the
 * array given to coh_scene_create (a GROUP_BEGIN index moves every member): the object becomes an
 * integer-pixel alias of itself (Cache.addtranslation); offsets accumulate. */
int coh_scene_translate_object(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy);
/* Render.shape_of_basicshape (render.ml:469-594) of the obj_index-th object (Path, Primitive or Group),
 * through the cache: entries are keyed by the object's id and hold the untranslated geometry's span
 * sets; the object's alias offset is applied on the way out (cache.ml:380-385). */
int coh_scene_object_shape(coh_ctx* ctx, coh_scene_t scene, int32_t obj_index, coh_shape_t* shape, coh_shape_t* minshape);
/* The covered-so-far set: export `u` as it stands after the scene pass of the last frame
 * rendered with COH_RENDER_RECORD_U (the set-subtraction artefact of render.ml:1308:
 * update minus every pixel the scene pass made opaque), as a device shape. */
int coh_render_uncovered(coh_ctx* ctx, coh_shape_t* out);
/* Multi-GPU band gather without a collective: peer_fbs are device pointers to the framebuffers of the other
 * GPUs of the box (same width x height, peer-mapped into this process: CUDA IPC / symmetric memory); from now on
 * every pixel this context renders into its own framebuffer is also stored to each of them over NVLink as it is
 * produced, so after all ranks have rendered their bands (and a cross-rank barrier) every GPU holds the whole
 * frame.  n_peers = 0 switches it off.  Frames with filter objects mirror only their final pixels. */
int coh_fb_set_peers(coh_ctx* ctx, int32_t n_peers, void* const* peer_fbs);
/* Bytes of device memory currently allocated from the stream-ordered pool every coh_* allocation comes from
 * (scenes, span sets, cache entries, framebuffer, scratch): for leak checks and for sizing the cache budget
 * that cache.mli:27-28 expresses in bytes. */
int coh_mem_in_use(coh_ctx* ctx, int64_t* bytes);
/* Wait for the context's stream and report deferred kernel-side failures. */
int coh_sync(coh_ctx* ctx);
/* Device pointer of the framebuffer (band gather by NCCL / peer copies happens on these). */
void* coh_fb_device_ptr(coh_ctx* ctx);
/* Copy a rectangle of the framebuffer to host memory: RGBA8, or the RGB888 layout that
 * Wxgui.plot_sprite writes (wxgui.ml:417-424, premultiplied r,g,b bytes, no alpha). */
int coh_fb_read_rgba(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out);
int coh_fb_read_rgb888(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out);
/* The value of Render.render_frame (render.mli:211-217: a Sprite.sprite on the update shape): the framebuffer's
 * pixels on `update`, one RGBA8 word per pixel in canonical span order — with coh_shape_export (update) the OCaml
 * shim rebuilds the sprite (spans of Fill.Samples).  The shape must lie inside the framebuffer. */
int coh_fb_read_sprite(coh_ctx* ctx, coh_shape_t update, uint32_t* rgba_out, int64_t cap, int64_t* n_out);
/* Asynchronous variant of coh_fb_read_rgba for a stream of frames (the refresh loop of wxgui.ml:333-367 reads
 * frame k while the engine already works on frame k+1): the rectangle is snapshotted on the render stream into
 * one of two staging buffers and copied to `out` (pinned host memory) on a copy stream, so the next
 * coh_scene_create / coh_render_frame overlap the transfer.  `out` is valid after coh_fb_read_wait, which waits
 * for every outstanding asynchronous read.  At most two reads are in flight; a third waits for the oldest. */
int coh_fb_read_rgba_async(coh_ctx* ctx, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out);
int coh_fb_read_wait(coh_ctx* ctx);

/* ---- several GPUs of one box ----
 * A frame shards by horizontal scanline bands (every stage is a pure function of the edge lists and the row): every
 * device holds the whole scene and renders its band; the gather of the RGBA8 strips is fused into the rendering
 * kernels — framebuffers are peer-mapped over NVLink and every finished pixel is stored to all of them
 * (coh_fb_set_peers) — so after a frame every device holds the whole picture and no collective follows.
 *
 * (1) ONE host process (the OCaml engine is one): coh_multi_* owns one context per device, enables peer access, and
 *     issues every device's launches from a worker thread of its own. */
typedef struct coh_multi coh_multi;
int coh_multi_init(int32_t n_devices, const int32_t* device_ids /* NULL: 0 .. n-1 */, coh_multi** out);
int coh_multi_shutdown(coh_multi* m);
const char* coh_multi_last_error(coh_multi* m);   /* m may be NULL for init errors */
int coh_multi_device_count(coh_multi* m);
coh_ctx* coh_multi_ctx(coh_multi* m, int32_t i);  /* the i-th device's context, for every other coh_* call */
/* framebuffer of width x height on every device; band k = rows [cuts[k], cuts[k+1]) (cuts = NULL: equal bands) */
int coh_multi_configure(coh_multi* m, int32_t width, int32_t height, const int32_t* cuts /* n_devices + 1, or NULL */);
int coh_multi_scene_create(coh_multi* m, const coh_object* objs, int32_t n_objs, int32_t n_background,
                           const int32_t* edges, int32_t n_edges, const int32_t* points, int32_t n_points, coh_scene_t* out);
int coh_multi_scene_free(coh_multi* m, coh_scene_t scene);
int coh_multi_scene_translate_object(coh_multi* m, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy);
/* Render.render_frame, every device its band; returns when the launches are issued */
int coh_multi_render_frame(coh_multi* m, coh_scene_t scene, int32_t ux, int32_t uy, int32_t uw, int32_t uh, int32_t flags);
int coh_multi_sync(coh_multi* m);   /* every device done: each framebuffer now holds the whole frame */
int coh_multi_fb_read_rgba(coh_multi* m, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out);   /* from device 0 */
int coh_multi_fb_read_rgb888(coh_multi* m, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out);
/* (2) one process per GPU (MPI / torchrun style hosts): the framebuffer is allocated for export, its 64-byte CUDA IPC
 *     handle travels through whatever channel the host has, and every other process maps it and passes the pointers
 *     to coh_fb_set_peers.  The host provides the cross-process barrier after a frame. */
int coh_fb_alloc_shared(coh_ctx* ctx, uint8_t handle_out[64]);   /* after coh_fb_configure; replaces the framebuffer */
int coh_fb_open_peer(coh_ctx* ctx, const uint8_t handle[64], void** device_ptr_out);
/*     Frame signals for such hosts (no collective library in the data path): a shared framebuffer carries COH_SIGNAL_SLOTS
 *     frame counters behind its pixels.  coh_frame_signal sets counter `slot` of every target framebuffer (pointers from
 *     coh_fb_open_peer, or this context's own coh_fb_device_ptr) to `epoch` once everything issued so far on the context's
 *     stream — the frame's kernels with their peer stores — is complete; coh_frame_wait makes the context's stream wait, on
 *     the device, until the listed counters of its own shared framebuffer have reached `epoch` (it gives up after ~4 s and
 *     the next coh_sync reports it).  The display rank waits for its peers' "band landed" counters; the peers wait for its
 *     "frame consumed" counter before they store the next frame (bench.py). */
#define COH_SIGNAL_SLOTS 16
int coh_frame_signal(coh_ctx* ctx, int32_t n_targets, void* const* target_fbs, int32_t slot, int32_t epoch);
int coh_frame_wait(coh_ctx* ctx, int32_t n_slots, const int32_t* slots, int32_t epoch);

/* ---- host-side geometry preparation (CPU; the step before the raster path) ----
 * A path segment record is 9 doubles: kind (0 straight, 1 cubic bezier) then up to four points.
 * Polygon.edgelist_of_path for one subpath (polygon.ml:119-127, 262-287; Coord.sub_of_float):
 * returns the number of edges and writes min(n, cap) of them as int32 x0,y0,x1,y1. */
int64_t coh_host_edgelist_of_subpath(const double* segs, int32_t n_segs, int32_t* edges_out, int64_t cap);
/* Brush.points_of_brushstroke for one subpath, rounded as in brush.ml:172 (polygon.ml:143-218):
 * returns the number of stamp centres and writes min(n, cap) of them as int32 x,y in list order. */
/* ---- brush strokes outside a scene (brush.mli:20-27) ----
 * `brush`: a BRUSH object record (brush_radius, brush_opacity, winding = COH_BRUSH_*, fill); points: the stroke's rounded
 * stamp points (Brush.points_of_brushstroke, brush.ml:126-130, 172; coh_host_brush_points), x, y pairs.
 * coh_brush_shape   Brush.shape_of_brushstroke (brush.ml:135-173): the boxes around the stamp points (the minshape is null).
 * coh_brush_sprite  Brush.sprite_of_brushstroke stroke fill shape (brush.ml:176-222): RGBA8 of every pixel of `shape` in span
 *                   order.  A dummy brush gives white (the reference returns the white sprite of the stroke's WHOLE shape
 *                   whatever shape it is asked for: pass coh_brush_shape's result).
 * coh_brush_smear   Brush.smear sprite stroke (brush.ml:286-331): the sprite (shape + RGBA8 in span order; shape 0 = NullSprite)
 *                   fleshed out to the stroke's shape and smeared along smear_points (Brush.find_smear_directions' integer
 *                   points, coh_host_smear_points); the result is a sprite on *out_shape = shape ∪ coh_brush_shape. */
int coh_brush_shape(coh_ctx* ctx, const coh_object* brush, const int32_t* points, int32_t n_points, coh_shape_t* shape);
int coh_brush_sprite(coh_ctx* ctx, const coh_object* brush, const int32_t* points, int32_t n_points, coh_shape_t shape,
                     uint32_t* rgba_out, int64_t cap, int64_t* n_out);
int coh_brush_smear(coh_ctx* ctx, coh_shape_t shape, const uint32_t* rgba_in, const coh_object* brush, const int32_t* points, int32_t n_points,
                    const int32_t* smear_points, int32_t n_smear, coh_shape_t* out_shape, uint32_t* rgba_out, int64_t cap, int64_t* n_out);
/* N2 — the step in front of the raster path on the device: Polygon.edgelist_of_path (polygon.ml:83-127, 262-287;
 * coord.ml:47) for the segments of a path (records of 9 doubles: kind 0 straight / 1 bezier, then up to 4 points, already
 * transformed to device space): de Casteljau subdivision to curve_accuracy, sub-bin edges in the reference's order.
 * coh_edgelist_of_path copies the edges back (n_out = their number, whatever cap is); coh_shapeminshape_of_path keeps them
 * in HBM and scan-converts them there (= Polygon.shapeminshape_polygon of the path, polygon.ml:605). */
int coh_edgelist_of_path(coh_ctx* ctx, const double* segs, int32_t n_segs, int32_t* edges_out, int64_t cap, int64_t* n_out);
int coh_shapeminshape_of_path(coh_ctx* ctx, const double* segs, int32_t n_segs, int32_t winding, coh_shape_t* shape, coh_shape_t* minshape);
/* N2 — the stroker: Shapes.strokepath (shapes.ml:203-530; shapes.mli:30-41).  A stroke specification is the reference's
 * record (shapes.ml:166-171).  The path: 9-double segment records of all its subpaths in order, subpath_segs[k] of them in
 * subpath k (closed subpaths are stroked like open ones, as in the reference, which reads the segments only).
 * coh_host_strokepath           Shapes.strokepath_polygon on the host: the outline as closed subpaths of straight and bezier
 *                               segments (rails, joins, caps; the circle of a degenerate path with round caps) and its
 *                               winding rule.  Returns the number of outline segments (whatever the caps are), -1 where
 *                               the reference fails.  The joins are made in Pdfutil.pair_reduce's order.
 * coh_strokepath                Shapes.strokepath: that outline flattened on the device (k_flatten) and sorted by
 *                               Polygon.sort_edgelist_maxy_rev (stable): the edge list a StrokedPath object carries.
 * coh_shapeminshape_of_stroke   the same edges kept in HBM and scan-converted there with the outline's winding rule. */
#define COH_CAP_BUTT 0
#define COH_CAP_ROUND 1
#define COH_CAP_PROJECTING 2
#define COH_JOIN_ROUND 0
#define COH_JOIN_MITRED 1
#define COH_JOIN_BEVEL 2
typedef struct coh_strokespec {
  int32_t startcap, join, endcap, reserved;   /* COH_CAP_*, COH_JOIN_*, COH_CAP_* */
  double mitrelimit, linewidth;
} coh_strokespec;
int64_t coh_host_strokepath(const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                            double* segs_out, int64_t cap_segs, int32_t* subpath_segs_out, int32_t cap_subpaths,
                            int32_t* n_subpaths_out, int32_t* winding_out);
/* Shapes.bounds_stroke (shapes.ml:522-540; Polygon.bounds_polygon, polygon.ml:404-438): xmin, xmax, ymin, ymax of the stroke in
 * pixels — the path's pixel box grown by the reach of its caps and joins.  Returns -1 for a path without subpaths. */
int32_t coh_host_bounds_stroke(const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths, int32_t bounds_out[4]);
int coh_strokepath(coh_ctx* ctx, const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                   int32_t* edges_out, int64_t cap, int64_t* n_out, int32_t* winding_out);
int coh_shapeminshape_of_stroke(coh_ctx* ctx, const coh_strokespec* spec, const double* segs, const int32_t* subpath_segs, int32_t n_subpaths,
                                coh_shape_t* shape, coh_shape_t* minshape);
/* Brush.points_of_brushstroke_smear and the integer points of Brush.find_smear_directions (brush.ml:239-283) for all the
 * segments of a path in order: pieces at most 2 apart, start points truncated, consecutive duplicates dropped. */
int64_t coh_host_smear_points(const double* segs, int32_t n_segs, int32_t* points_out, int64_t cap);
int64_t coh_host_brush_points(const double* segs, int32_t n_segs, double radius, int32_t* points_out, int64_t cap);

/* ---- N4 meeting N1 — the front end's socket format (camlpy.mli; camlpy.ml:18-124, Python side pycaml.py:30-98) ----
 * A Camlpy.marshallable crosses the ABI as its pre-order token list: kinds[i] = COH_WIRE_* (the format's own tags,
 * camlpy.ml:26-30); values[i] = the Int (written as its low 32 bits, read back as 0 .. 2^32 - 1 without sign extension,
 * camlpy.ml:33-37, 85-86), the Bool (0 / 1; any non-zero byte reads as 1), the String's length, or the number of members of
 * the Tuple; offsets[i] = where a String's bytes lie — in `strings` for marshal, in the message for unmarshal.
 * coh_host_wire_marshal     Camlpy.marshall: 4 bytes of size + the flattened value.  Returns the message size (written only
 *                           when out != NULL and cap suffices), -1 for a token list that is not exactly one value.
 * coh_host_wire_unmarshal   Camlpy.unmarshall: *taken = 0 while the message is incomplete (None), else the bytes taken
 *                           (Some (len + 4, v)); returns -1 for Invalid_data (malformed, or not exactly one value).
 *                           *n_tokens is the number of tokens whatever cap_tokens is.
 * coh_host_wire_refresh_window   the bytes in front of the pixels of Wxgui.refresh_window's message (wxgui.ml:352-366):
 *                           Tuple [String "RefreshWindow"; Int window; Int xmin; Int ymin; Int w; Int h; String rgb888].
 *                           Returns the size of the whole message, 0 where the reference sends nothing (xmin = xmax or
 *                           ymin = ymax), -1 where its assertion fails (wxgui.ml:335).
 * coh_wire_refresh_window   the whole message, its pixel string (string_of_canvas_portion, wxgui.ml:334-350: rows ymin ..
 *                           ymax, columns xmin .. xmax INCLUSIVE, r, g, b) copied out of the GPU framebuffer; *len = its
 *                           size (0: nothing to send); nothing is written when cap < *len.  The reference's canvas is
 *                           1280 x 1024 "for now" (wxgui.ml:336); here the rectangle must lie inside the framebuffer. */
#define COH_WIRE_TUPLE 0
#define COH_WIRE_UNIT 1
#define COH_WIRE_INT 2
#define COH_WIRE_STRING 3
#define COH_WIRE_BOOL 4
int64_t coh_host_wire_marshal(const int32_t* kinds, const int64_t* values, const int64_t* offsets, int32_t n_tokens,
                              const uint8_t* strings, uint8_t* out, int64_t cap);
int32_t coh_host_wire_unmarshal(const uint8_t* buf, int64_t n, int32_t* kinds, int64_t* values, int64_t* offsets, int32_t cap_tokens,
                                int32_t* n_tokens, int64_t* taken);
int64_t coh_host_wire_refresh_window(int32_t window, int32_t xmin, int32_t ymin, int32_t xmax, int32_t ymax, uint8_t header_out[64], int32_t* header_len);
int coh_wire_refresh_window(coh_ctx* ctx, int32_t window, int32_t xmin, int32_t ymin, int32_t xmax, int32_t ymax, uint8_t* out, int64_t cap, int64_t* len);

#ifdef __cplusplus
}
#endif
#endif /* COHERENCE_B200_H */
