(* coherence_gpu.ml — the OCaml side of the drop-in boundary (see INTEGRATION.md).
   NOT COMPILED HERE (no OCaml toolchain in this image or on the GPU box); reviewed source, written against the
   reference's own .mli files (render.mli, sprite.mli, fill.mli, polygon.mli, brush.mli, convolve.mli, id.mli).

   The reference's entry points keep their signatures; their bodies call the externals below (one per symbol of
   include/coherence_b200.h).  What cannot cross a C ABI is mapped here:
     - Fill.fill is a record of closures (fill.mli:25-30).  A Plain fill's colour is [fillsingle 0 0], as the
       reference itself does (sprite.ml:171, engine.ml:315-316).  Gradient and radial fills are recognised through a
       side table: build them with [gradient] / [radial] below (they call Fill.gradient / Fill.radial and remember the
       parameters under the physical identity of the record); [filltransform] results are registered the same way.
     - Brush.brush is abstract (brush.mli:6): brushstrokes enter through [register_brush].
     - Render.filter holds closures: the filters of filters.ml enter through [register_filter].
     - geometry crosses AFTER the transform: Polygon.transform_path / edgelist_of_path run here, on the host, exactly
       as in the reference (render.ml:197-205, polygon.ml:284-287); no arithmetic of camlpdf is on the device path. *)
open Bigarray

type ctx = nativeint
type shape_h = int64   (* device-resident Sprite.shape; 0L = NullShape *)
type scene_h = int64
type i32 = (int32, int32_elt, c_layout) Array1.t
type u8 = (int, int8_unsigned_elt, c_layout) Array1.t
type f64 = (float, float64_elt, c_layout) Array1.t
type i64 = (int64, int64_elt, c_layout) Array1.t

(* ---- one external per exported symbol (ocaml/coherence_stubs.c) ---- *)
external init : int -> ctx = "coh_ml_init"
external shutdown : ctx -> unit = "coh_ml_shutdown"
external device_name : ctx -> string = "coh_ml_device_name"
external stream : ctx -> nativeint = "coh_ml_stream"
external set_stream : ctx -> nativeint -> unit = "coh_ml_set_stream"
external launch_count : ctx -> int64 = "coh_ml_launch_count"
external set_timing : ctx -> bool -> unit = "coh_ml_set_timing"
external get_timing : ctx -> float * float * int = "coh_ml_get_timing"
external set_option : ctx -> string -> int -> unit = "coh_ml_set_option"
external mem_in_use : ctx -> int64 = "coh_ml_mem_in_use"
external sync : ctx -> unit = "coh_ml_sync"
external rgba8_of_colour : Colour.colour -> int32 = "coh_ml_rgba8_of_colour"
external colour_of_rgba8 : int32 -> Colour.colour = "coh_ml_colour_of_rgba8"
external shapeminshape : ctx -> i32 -> int -> shape_h * shape_h = "coh_ml_shapeminshape"
(* Brush.shape_of_brushstroke / sprite_of_brushstroke / smear outside a scene (brush.mli:20-27): the brush as a packed BRUSH
   record, the stroke as its rounded stamp points *)
external brush_shape : ctx -> u8 -> i32 -> shape_h = "coh_ml_brush_shape"
external brush_sprite : ctx -> u8 -> i32 -> shape_h -> i32 -> int = "coh_ml_brush_sprite"
external brush_smear : ctx -> shape_h -> i32 -> u8 -> i32 -> i32 -> i32 -> shape_h * int = "coh_ml_brush_smear_bc" "coh_ml_brush_smear"
(* N2: Polygon.edgelist_of_path / shapeminshape_polygon with the flattening on the device (segments as 9 floats each) *)
external edgelist_of_path : ctx -> f64 -> i32 -> int = "coh_ml_edgelist_of_path"
external shapeminshape_of_path : ctx -> f64 -> int -> shape_h * shape_h = "coh_ml_shapeminshape_of_path"
(* N2: Shapes.strokepath — the outline by the library's host stroker, flattened (and scan-converted) on the device.
   spec = [| startcap; join; endcap; mitrelimit; linewidth |], counts = the segments of each subpath *)
external strokepath_raw : ctx -> f64 -> f64 -> i32 -> i32 -> int * int = "coh_ml_strokepath"
external shapeminshape_of_stroke : ctx -> f64 -> f64 -> i32 -> shape_h * shape_h = "coh_ml_shapeminshape_of_stroke"
external host_strokepath : f64 -> f64 -> i32 -> f64 -> i32 -> int * int * int = "coh_ml_host_strokepath"
external host_bounds_stroke : f64 -> f64 -> i32 -> int * int * int * int = "coh_ml_host_bounds_stroke"
external host_wire_marshal : i32 -> i64 -> i64 -> string -> string = "coh_ml_host_wire_marshal"
external host_wire_unmarshal : string -> i32 -> i64 -> i64 -> int * int = "coh_ml_host_wire_unmarshal"
external host_wire_refresh_window : int -> int * int * int * int -> string * int = "coh_ml_host_wire_refresh_window"
external wire_refresh_window : ctx -> int -> int * int * int * int -> string = "coh_ml_wire_refresh_window"
external polygon_opacity : ctx -> i32 -> int -> shape_h -> u8 -> int = "coh_ml_polygon_opacity"
external polygon_sprite_raw : ctx -> u8 -> i32 -> int -> shape_h -> i32 -> int = "coh_ml_polygon_sprite_bc" "coh_ml_polygon_sprite"
external shape_box : ctx -> int -> int -> int -> int -> shape_h = "coh_ml_shape_box"
external shape_export : ctx -> shape_h -> i32 = "coh_ml_shape_export"
external shape_import : ctx -> i32 -> shape_h = "coh_ml_shape_import"
external shape_bounds : ctx -> shape_h -> (int * int * int * int) option = "coh_ml_shape_bounds"
external shape_card : ctx -> shape_h -> int = "coh_ml_shape_card"
external shape_free : ctx -> shape_h -> unit = "coh_ml_shape_free"
external shape_union : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_union"
external shape_difference : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_difference"
external shape_intersection : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_intersection"
external shape_translate : ctx -> shape_h -> int -> int -> shape_h = "coh_ml_shape_translate"
external shape_bloat : ctx -> shape_h -> int -> int -> shape_h = "coh_ml_shape_bloat"
external shape_erode : ctx -> shape_h -> int -> int -> shape_h = "coh_ml_shape_erode"
external shape_intersects_raw : ctx -> shape_h -> shape_h -> bool = "coh_ml_shape_intersects"
external sprite_portion_raw : ctx -> shape_h -> i32 -> shape_h -> i32 -> int = "coh_ml_sprite_portion"
external sprite_fillshape_raw : ctx -> shape_h -> u8 -> i32 -> int = "coh_ml_sprite_fillshape"
external sprite_map_raw : ctx -> int * int -> i32 -> i32 -> unit = "coh_ml_sprite_map"
external sprite_map_coords_fill_raw : ctx -> shape_h -> u8 -> i32 -> i32 -> int = "coh_ml_sprite_map_coords_fill"
external convolve_sprite_raw : ctx -> int * int -> shape_h -> i32 -> i32 -> shape_h = "coh_ml_convolve_sprite"
external cache_configure : ctx -> bool -> int64 -> unit = "coh_ml_cache_configure"
external cache_clear : ctx -> unit = "coh_ml_cache_clear"
external cache_stats : ctx -> int * int * int * int = "coh_ml_cache_stats"
external cache_sprite_stats : ctx -> scene_h -> int * int * int * int = "coh_ml_cache_sprite_stats"
external cache_addshape : ctx -> int64 -> shape_h -> shape_h -> unit = "coh_ml_cache_addshape"
external cache_getshape : ctx -> int64 -> (shape_h * shape_h) option = "coh_ml_cache_getshape"
external cache_addtranslation : ctx -> int64 -> int64 -> int -> int -> unit = "coh_ml_cache_addtranslation"
external dirty_region_raw : ctx -> shape_h * shape_h * shape_h * shape_h -> shape_h -> bool -> shape_h = "coh_ml_dirty_region"
external pack_object : u8 -> int -> int array -> float array -> int64 -> unit = "coh_ml_pack_object"
external sizeof_object : unit -> int = "coh_ml_sizeof_object"
external scene_create : ctx -> u8 -> int -> i32 -> i32 -> scene_h = "coh_ml_scene_create"
external scene_free : ctx -> scene_h -> unit = "coh_ml_scene_free"
external fb_configure : ctx -> int -> int -> int -> int -> unit = "coh_ml_fb_configure"
external fb_attach : ctx -> nativeint -> unit = "coh_ml_fb_attach"
external fb_device_ptr : ctx -> nativeint = "coh_ml_fb_device_ptr"
external fb_set_peers : ctx -> nativeint array -> unit = "coh_ml_fb_set_peers"
external render_frame_box : ctx -> scene_h -> int * int * int * int -> int -> unit = "coh_ml_render_frame"
external render_frame_rgb888 : ctx -> scene_h -> int * int * int * int -> u8 -> unit = "coh_ml_render_frame_rgb888"
external render_frame_shape : ctx -> scene_h -> shape_h -> int -> unit = "coh_ml_render_frame_shape"
external render_uncovered : ctx -> shape_h = "coh_ml_render_uncovered"
external fb_read_sprite : ctx -> shape_h -> i32 = "coh_ml_fb_read_sprite"
external read_rgba : ctx -> int * int * int * int -> u8 -> unit = "coh_ml_read_rgba"
external read_rgba_async : ctx -> int * int * int * int -> u8 -> unit = "coh_ml_read_rgba_async"
external read_wait : ctx -> unit = "coh_ml_read_wait"
external read_rgb888 : ctx -> int * int * int * int -> u8 -> unit = "coh_ml_read_rgb888"
external scene_translate_object : ctx -> scene_h -> int -> int -> int -> unit = "coh_ml_scene_translate_object"
external scene_drag_object : ctx -> scene_h -> int -> int -> int -> int * int * int * int = "coh_ml_scene_drag_object"
external scene_object_shape : ctx -> scene_h -> int -> shape_h * shape_h = "coh_ml_scene_object_shape"
external dirty_filter : ctx -> scene_h -> int -> shape_h -> shape_h = "coh_ml_dirty_filter"
type multi = nativeint
external multi_init : int array -> multi = "coh_ml_multi_init"
external multi_shutdown : multi -> unit = "coh_ml_multi_shutdown"
external multi_device_count : multi -> int = "coh_ml_multi_device_count"
external multi_ctx : multi -> int -> ctx = "coh_ml_multi_ctx"
external multi_configure : multi -> int -> int -> unit = "coh_ml_multi_configure"
external multi_scene_create : multi -> u8 -> int -> i32 -> i32 -> scene_h = "coh_ml_multi_scene_create"
external multi_scene_free : multi -> scene_h -> unit = "coh_ml_multi_scene_free"
external multi_scene_translate_object : multi -> scene_h -> int -> int -> int -> unit = "coh_ml_multi_scene_translate_object"
external multi_render_frame : multi -> scene_h -> int * int * int * int -> int -> unit = "coh_ml_multi_render_frame"
external multi_sync : multi -> unit = "coh_ml_multi_sync"
external multi_read_rgba : multi -> int * int * int * int -> u8 -> unit = "coh_ml_multi_read_rgba"
external multi_read_rgb888 : multi -> int * int * int * int -> u8 -> unit = "coh_ml_multi_read_rgb888"
external fb_alloc_shared : ctx -> u8 -> unit = "coh_ml_fb_alloc_shared"
external fb_open_peer : ctx -> u8 -> nativeint = "coh_ml_fb_open_peer"
(* frame counters behind the pixels of a shared framebuffer: "band landed" / "frame consumed" between processes *)
external frame_signal : ctx -> nativeint array -> int -> int -> unit = "coh_ml_frame_signal"
external frame_wait : ctx -> i32 -> int -> unit = "coh_ml_frame_wait"
external host_edgelist_of_subpath : f64 -> i32 -> int = "coh_ml_host_edgelist_of_subpath"
external host_brush_points : f64 -> float -> i32 -> int = "coh_ml_host_brush_points"
external host_smear_points : f64 -> i32 -> int = "coh_ml_host_smear_points"

let the_ctx = lazy (init (-1))
let ctx () = Lazy.force the_ctx

(* the canvas of wxgui.ml:254-262 is 1280 x 1024; the device framebuffer takes its place *)
let canvas = ref (1280, 1024)
let set_canvas w h = canvas := (w, h); fb_configure (ctx ()) w h 0 h
let canvas_ready = ref false
let ensure_canvas () = if not !canvas_ready then begin let w, h = !canvas in set_canvas w h; canvas_ready := true end

(* ---- Sprite.shape <-> flat records (y, nspans, (x, len) ...), rows grouped into vspans by consecutive y ---- *)
let rows_of_flat (a : i32) : (int * (int * int) list) list =
  let n = Array1.dim a in
  let rows = ref [] and i = ref 0 in
  while !i < n do
    let y = Int32.to_int a.{!i} and k = Int32.to_int a.{!i + 1} in
    let spans = List.init k (fun q -> (Int32.to_int a.{!i + 2 + 2 * q}, Int32.to_int a.{!i + 3 + 2 * q})) in
    rows := (y, spans) :: !rows;
    i := !i + 2 + 2 * k
  done;
  List.rev !rows   (* increasing y *)

let vspans_of_rows rows =
  (* maximal runs of consecutive y (sprite.ml:631-634 vspan_accumulate), in increasing order *)
  let acc =
    List.fold_left
      (fun acc (y, line) ->
         match acc with
         | (s, l, lines) :: rest when y = s + l -> (s, l + 1, line :: lines) :: rest
         | _ -> (y, 1, [line]) :: acc)
      [] rows
  in
  List.rev_map (fun (s, l, lines) -> (s, l, List.rev lines)) acc

let shape_of_flat (a : i32) : Sprite.shape =
  if Array1.dim a = 0 then Sprite.NullShape
  else Sprite.boxshape (Sprite.Shape (Sprite.NoBounds, vspans_of_rows (rows_of_flat a)))

let flat_of_shape (s : Sprite.shape) : i32 =
  match s with
  | Sprite.NullShape -> Array1.create int32 c_layout 0
  | Sprite.Shape (_, vspans) ->
    let buf = ref [] and n = ref 0 in
    let push v = buf := Int32.of_int v :: !buf; incr n in
    List.iter
      (fun (s0, _, lines) ->
         List.iteri (fun k line -> push (s0 + k); push (List.length line); List.iter (fun (x, l) -> push x; push l) line) lines)
      vspans;
    let a = Array1.create int32 c_layout !n in
    List.iteri (fun i v -> a.{!n - 1 - i} <- v) !buf;
    a

(* a device copy of a Sprite.shape for the duration of [f] *)
let with_shape (s : Sprite.shape) f =
  let h = shape_import (ctx ()) (flat_of_shape s) in
  match f h with
  | r -> shape_free (ctx ()) h; r
  | exception e -> shape_free (ctx ()) h; raise e

let take_shape h = let s = shape_of_flat (shape_export (ctx ()) h) in shape_free (ctx ()) h; s

(* (shape, RGBA8 per pixel in span order) -> Sprite.sprite: every span one subspan of Fill.Samples *)
let sprite_of_pixels (shape : i32) (px : i32) : Sprite.sprite =
  if Array1.dim shape = 0 then Sprite.NullSprite
  else begin
    let k = ref 0 in
    let rows =
      List.map
        (fun (y, spans) ->
           (y, List.map
              (fun (x, l) ->
                 let arr = Array.init l (fun i -> colour_of_rgba8 px.{!k + i}) in
                 k := !k + l;
                 (x, l, [(l, Fill.Samples arr)]))
              spans))
        (rows_of_flat shape)
    in
    (* the sprite's bounds are those of its shape (Sprite.boxshape, sprite.ml:542-549) *)
    let bounds = match shape_of_flat shape with Sprite.Shape (b, _) -> b | Sprite.NullShape -> Sprite.NoBounds in
    Sprite.Sprite (bounds, vspans_of_rows rows)
  end

(* ---- polygon.mli:44-59 ---- *)
let int_of_winding = function Pdfgraphics.NonZero -> 0 | Pdfgraphics.EvenOdd -> 1

let i32_of_edges (edges : Polygon.edge list) : i32 =
  let a = Array1.create int32 c_layout (4 * List.length edges) in
  List.iteri
    (fun i (e : Polygon.edge) ->
       a.{4 * i} <- Int32.of_int e.Polygon.x0; a.{4 * i + 1} <- Int32.of_int e.Polygon.y0;
       a.{4 * i + 2} <- Int32.of_int e.Polygon.x1; a.{4 * i + 3} <- Int32.of_int e.Polygon.y1)
    edges;
  a

let shapeminshape_of_unsorted_edgelist (edges : Polygon.edge list) winding =
  let s, m = shapeminshape (ctx ()) (i32_of_edges edges) (int_of_winding winding) in
  let r = take_shape s in
  (r, take_shape m)

let shapeminshape_polygon (path : Pdfgraphics.path) =
  shapeminshape_of_unsorted_edgelist (Polygon.edgelist_of_path path) (fst path)

(* ---- shapes.mli:38-41: Shapes.strokepath / bounds_stroke through the library (N2) ---- *)
let f64_of_strokespec (sp : Shapes.strokespec) : f64 =
  let cap = function Shapes.ButtCap -> 0. | Shapes.RoundCap -> 1. | Shapes.ProjectingCap -> 2.
  and join = function Shapes.RoundJoin -> 0. | Shapes.MitredJoin -> 1. | Shapes.BevelJoin -> 2. in
  let a = Array1.create float64 c_layout 5 in
  a.{0} <- cap sp.Shapes.startcap; a.{1} <- join sp.Shapes.join; a.{2} <- cap sp.Shapes.endcap;
  a.{3} <- sp.Shapes.mitrelimit; a.{4} <- sp.Shapes.linewidth;
  a

(* the segments of every subpath as 9-float records (kind, then up to four points) + the segments per subpath *)
let records_of_path ((_, subpaths) : Pdfgraphics.path) : f64 * i32 =
  let segs = List.concat (List.map (fun (_, _, s) -> s) subpaths) in
  let rec_ = Array1.create float64 c_layout (9 * max 1 (List.length segs)) in
  Array1.fill rec_ 0.;
  List.iteri
    (fun i seg ->
       let put k (x, y) = rec_.{9 * i + 1 + 2 * k} <- x; rec_.{9 * i + 2 + 2 * k} <- y in
       match seg with
       | Pdfgraphics.Straight (a, b) -> rec_.{9 * i} <- 0.; put 0 a; put 1 b
       | Pdfgraphics.Bezier (a, b, c, d) -> rec_.{9 * i} <- 1.; put 0 a; put 1 b; put 2 c; put 3 d)
    segs;
  let counts = Array1.create int32 c_layout (List.length subpaths) in
  List.iteri (fun k (_, _, s) -> counts.{k} <- Int32.of_int (List.length s)) subpaths;
  (Array1.sub rec_ 0 (9 * List.length segs), counts)

(* val strokepath : strokespec -> Pdfgraphics.path -> Polygon.edge list — the outline by the library's host stroker,
   flattened on the device, sorted by Polygon.sort_edgelist_maxy_rev *)
let strokepath (sp : Shapes.strokespec) (path : Pdfgraphics.path) : Polygon.edge list =
  let spec = f64_of_strokespec sp and segs, counts = records_of_path path in
  let rec go cap =
    let out = Array1.create int32 c_layout (4 * cap) in
    let n, _ = strokepath_raw (ctx ()) spec segs counts out in
    if n > cap then go n
    else List.init n (fun i ->
        { Polygon.x0 = Int32.to_int out.{4 * i}; Polygon.y0 = Int32.to_int out.{4 * i + 1};
          Polygon.x1 = Int32.to_int out.{4 * i + 2}; Polygon.y1 = Int32.to_int out.{4 * i + 3} })
  in
  go (256 * (Array1.dim segs / 9) + 256)

(* val bounds_stroke : Pdfgraphics.path -> strokespec -> int * int * int * int *)
let bounds_stroke (path : Pdfgraphics.path) (sp : Shapes.strokespec) =
  let segs, counts = records_of_path path in
  host_bounds_stroke (f64_of_strokespec sp) segs counts

(* ---- fills: descriptors for the records of closures (fill.mli) ---- *)
type fill_desc =
  | FPlain of Colour.colour
  | FAxial of (float * float) * (float * float) * bool * bool * Colour.colour * Colour.colour
  | FRadial of (float * float) * (float * float) * (float * float) * bool * bool * Colour.colour * Colour.colour

let fill_table : (Fill.fill * fill_desc) list ref = ref []   (* physical identity; a Weak table in production *)
let remember f d = fill_table := (f, d) :: !fill_table; f
let gradient p0 p1 es ee cs ce = remember (Fill.gradient p0 p1 es ee cs ce) (FAxial (p0, p1, es, ee, cs, ce))
let radial c p p' es ee cs ce = remember (Fill.radial c p p' es ee cs ce) (FRadial (c, p, p', es, ee, cs, ce))

let desc_of_fill (f : Fill.fill) : fill_desc =
  match f.Fill.fillkind with
  | Fill.Plain -> FPlain (f.Fill.fillsingle 0 0)
  | Fill.Fancy ->
    (try List.assq f !fill_table
     with Not_found -> failwith "Coherence_gpu: a Fancy fill that was not built with Coherence_gpu.gradient / radial")

(* Fill.filltransform tr fill (render.ml:1008): the descriptor's points through the same transform *)
let transform_desc tr = function
  | FPlain c -> FPlain c
  | FAxial (p0, p1, es, ee, cs, ce) -> FAxial (Pdftransform.transform tr p0, Pdftransform.transform tr p1, es, ee, cs, ce)
  | FRadial (c, p, p', es, ee, cs, ce) ->
    FRadial (Pdftransform.transform tr c, Pdftransform.transform tr p, Pdftransform.transform tr p', es, ee, cs, ce)

(* ---- brushes and filters registered by the caller (abstract type / closures) ---- *)
let brush_table : (Brush.brushstroke * (float * float)) list ref = ref []   (* (radius, opacity) of Brush.mkround *)
let register_brush (b : Brush.brushstroke) ~radius ~opacity = brush_table := (b, (radius, opacity)) :: !brush_table; b

(* Convolve.kernel is abstract (convolve.mli:4): build kernels with [mkunit] / [mkgaussian] below *)
let kernel_table : (Convolve.kernel * int) list ref = ref []   (* 1 = UnitKernel, 2 = XYKernel of mkgaussian *)
let mkunit r = let k = Convolve.mkunit r in kernel_table := (k, 1) :: !kernel_table; k
let mkgaussian r = let k = Convolve.mkgaussian r in kernel_table := (k, 2) :: !kernel_table; k
let kind_of_kernel k =
  try List.assq k !kernel_table
  with Not_found -> failwith "Coherence_gpu: a kernel that was not built with Coherence_gpu.mkunit / mkgaussian"

type filter_desc =
  | Hole | Monochrome
  | Blur of Convolve.kernel * int          (* the kernel and its radius *)
  | Rewritten of (Render.scene -> Render.scene)   (* affine / rgb / wireframe / swapdepth: the reading scene *)
let filter_table : (Render.filter * filter_desc) list ref = ref []
let register_filter (f : Render.filter) d = filter_table := (f, d) :: !filter_table; f

(* ---- flattening a scene into the ABI's arrays (render.ml:19-75 -> coh_object records) ---- *)
type flat = {
  mutable recs : (int array * float array * int64) list;   (* reversed *)
  mutable n_recs : int;
  mutable edges : Polygon.edge list list;                   (* reversed chunks *)
  mutable n_edges : int;
  mutable points : (int * int) list list;
  mutable n_points : int;
  mutable n_background : int;
  mutable reading : (int * Render.scene) list;              (* filter record index, its rewritten scene *)
}
let new_flat () = { recs = []; n_recs = 0; edges = []; n_edges = 0; points = []; n_points = 0; n_background = 0; reading = [] }

(* field order of coh_ml_pack_object *)
let blank kind =
  let ints = Array.make 28 0 in
  ints.(0) <- kind; ints.(8) <- -1;
  (ints, Array.make 8 0., -1L)

let k_path = 0 and k_primitive = 1 and k_group_begin = 2 and k_group_end = 3 and k_brush = 4 and k_cpg = 5 and k_filter = 6

let push fl r = fl.recs <- r :: fl.recs; fl.n_recs <- fl.n_recs + 1
let put_edges fl (e : Polygon.edge list) =
  let first = fl.n_edges in
  fl.edges <- e :: fl.edges; fl.n_edges <- fl.n_edges + List.length e; first

let set_fill (ints, floats, _) d =
  let c w = Int32.to_int (rgba8_of_colour w) land 0xFFFFFFFF in
  match d with
  | FPlain col -> ints.(4) <- 0; ints.(5) <- c col
  | FAxial ((x0, y0), (x1, y1), es, ee, cs, ce) ->
    ints.(4) <- 1; ints.(5) <- c cs; ints.(6) <- c ce; ints.(7) <- (if es then 1 else 0) lor (if ee then 2 else 0);
    floats.(0) <- x0; floats.(1) <- y0; floats.(2) <- x1; floats.(3) <- y1
  | FRadial ((cx, cy), (px, py), (qx, qy), es, ee, cs, ce) ->
    ints.(4) <- 2; ints.(5) <- c cs; ints.(6) <- c ce; ints.(7) <- (if es then 1 else 0) lor (if ee then 2 else 0);
    floats.(0) <- cx; floats.(1) <- cy; floats.(2) <- px; floats.(3) <- py; floats.(4) <- qx; floats.(5) <- qy

let pretrans_of_compop = function
  | Render.Over | Render.NoCover -> -1
  | Render.PreTrans (v, _) -> int_of_float (v *. 255.)      (* render.ml:1295-1298 toint (v *. 255.) *)

(* Id.idset = id * hash: the hash is the cache key on the device (cache.ml:82-83 keys its table the same way) *)
let key_of_idset (ids : Id.idset) : int64 = Int64.of_int (snd ids)

let cpg_code = function Render.Union -> 0 | Render.Intersection -> 1 | Render.Subtraction -> 2 | Render.ExclusiveOr -> 3
let kernel_code k r = kind_of_kernel k lor (r lsl 8)

(* transform_shapekind (render.ml:197-205) through the exported Render.transform_basicshape (render.mli:158) *)
let transform_shapekind tr sk =
  match Render.transform_basicshape tr (Render.Basic (Fill.dummy, sk)) with
  | Render.Basic (_, sk') -> sk'
  | _ -> sk

let edges_of_basic tr sk =
  match transform_shapekind tr sk with
  | Render.Path p -> (Polygon.edgelist_of_path p, int_of_winding (fst p), 0)
  | Render.StrokedPath (p, spec) ->
    (* shape by NonZero (render.ml:510), sprite by EvenOdd (render.ml:1018) *)
    (Shapes.strokepath spec p, 0, 1 + 1)
  | _ -> failwith "Coherence_gpu: CPG operands must be paths"

let rec flatten_obj fl ?(cached = true) ?(conv = 0) (Render.Obj (ids, geom, tr, compop)) =
  let id = if cached then key_of_idset ids else -1L in
  match geom with
  | Render.Group objs ->
    (* render.ml:988-1001: the group's transform is appended onto its members; members get fresh ids (never cached).
       conv <> 0: Convolved (k, Group objs) — the GROUP_BEGIN record carries the kernel (coh_object.convolve) *)
    let (ints, floats, _) = blank k_group_begin in
    ints.(8) <- pretrans_of_compop compop; ints.(20) <- conv;
    push fl (ints, floats, id);
    List.iter (fun (Render.Obj (i, g, tr', c)) -> flatten_obj fl ~cached:false (Render.Obj (i, g, Pdftransform.append tr tr', c))) objs;
    push fl (blank k_group_end)
  | Render.Primitive (col, p) ->
    let (ints, floats, _) = blank k_primitive in
    ints.(8) <- pretrans_of_compop compop;
    ints.(5) <- Int32.to_int (rgba8_of_colour col) land 0xFFFFFFFF;
    let ti = int_of_float in
    ignore tr;   (* shape_of_basicshape's Primitive branch does not apply the object's transform (render.ml:556-586) *)
    let (x0, y0, x1, y1, null) =
      match p with
      | Render.HLine (y, xmin, xmax) -> (ti xmin, ti y, ti xmax, ti y, ti xmax = ti xmin)          (* render.ml:558-565 *)
      | Render.VLine (x, ymin, ymax) -> (ti x, ti ymin, ti x, ti ymax, ti ymax = ti ymin)          (* render.ml:566-572 *)
      | Render.Rectangle (xmin, ymin, xmax, ymax) -> (ti xmin, ti ymin, ti xmax, ti ymax, false)   (* render.ml:573-586 *)
    in
    ints.(15) <- x0; ints.(16) <- y0; ints.(17) <- x1; ints.(18) <- y1; ints.(19) <- (if null then 1 else 0);
    ints.(11) <- x0; ints.(12) <- x1; ints.(13) <- y0; ints.(14) <- y1;
    push fl (ints, floats, id)
  | Render.Basic (fill, shapekind) -> flatten_basic fl id compop tr (transform_desc tr (desc_of_fill fill)) shapekind 0
  | Render.Convolved (k, Render.Basic (fill, shapekind)) ->
    flatten_basic fl id compop tr (transform_desc tr (desc_of_fill fill)) shapekind (kernel_code k (Convolve.radius_of_kernel k))
  | Render.Convolved (k, (Render.Group _ as g)) ->
    (* render.ml:1023-1052 with a Group child: members with fancy fills are refused by coh_scene_create (DESIGN.md) *)
    flatten_obj fl ~cached ~conv:(kernel_code k (Convolve.radius_of_kernel k)) (Render.Obj (ids, g, tr, compop))
  | Render.Convolved (_, _) -> failwith "Coherence_gpu: Convolved of this geometry (Convolved, Filter, Primitive) is not handled by the device path"
  | Render.Filter f ->
    let d = try List.assq f !filter_table with Not_found -> failwith "Coherence_gpu: a filter that was not registered (register_filter)" in
    (match f.Render.geometry with
     | Render.Basic (fill, Render.Path p) ->
       let p' = Polygon.transform_path tr p in
       let (ints, floats, _) as r = blank k_filter in
       ints.(2) <- put_edges fl (Polygon.edgelist_of_path p'); ints.(3) <- fl.n_edges - ints.(2); ints.(1) <- int_of_winding (fst p');
       set_fill r (transform_desc tr (desc_of_fill fill));
       (match d with
        | Hole -> ints.(26) <- 1
        | Monochrome -> ints.(26) <- 2
        | Blur (k, r') -> ints.(26) <- 3; ints.(27) <- kernel_code k r'
        | Rewritten rewrite -> ints.(26) <- 4; fl.reading <- (fl.n_recs, rewrite []) :: fl.reading);
       ignore floats;
       push fl (ints, floats, -1L)
     | _ -> failwith "Coherence_gpu: filter geometry must be Basic (fill, Path _)")

and flatten_basic fl id compop tr fdesc shapekind conv =
  let (ints, floats, _) as r = blank k_path in
  ints.(8) <- pretrans_of_compop compop; ints.(20) <- conv;
  set_fill r fdesc;
  (match shapekind with
   | Render.Path _ | Render.StrokedPath _ ->
     let (e, w, sw) = edges_of_basic tr shapekind in
     ints.(2) <- put_edges fl e; ints.(3) <- List.length e; ints.(1) <- w; ints.(21) <- sw;
     (* bounds_of_basicshape (render.ml:377-437) = pix_of_sub of the edge extremes *)
     if e <> [] then begin
       let xs = List.concat_map (fun (d : Polygon.edge) -> [d.Polygon.x0; d.Polygon.x1]) e
       and ys = List.concat_map (fun (d : Polygon.edge) -> [d.Polygon.y0; d.Polygon.y1]) e in
       let mn = List.fold_left min max_int and mx = List.fold_left max min_int in
       ints.(11) <- Coord.pix_of_sub (mn xs); ints.(12) <- Coord.pix_of_sub (mx xs);
       ints.(13) <- Coord.pix_of_sub (mn ys); ints.(14) <- Coord.pix_of_sub (mx ys)
     end
   | Render.CPG (op, a, b) ->
     let (ea, wa, _) = edges_of_basic tr a and (eb, wb, _) = edges_of_basic tr b in
     ints.(0) <- k_cpg;
     ints.(2) <- put_edges fl ea; ints.(3) <- List.length ea; ints.(1) <- wa;
     ints.(22) <- put_edges fl eb; ints.(23) <- List.length eb; ints.(24) <- wb; ints.(25) <- cpg_code op
   | Render.Brushstroke bs ->
     let (radius, opacity) =
       try List.assq bs !brush_table with Not_found -> failwith "Coherence_gpu: a brushstroke that was not registered (register_brush)" in
     let bs' = Brush.transform_brushstroke tr bs in   (* render.ml:200-201 *)
     (* Brush.points_of_brushstroke (brush.ml:126-130, 172): Polygon.points_on_path at w / 20, rounded toint (v +. 0.5) *)
     let w = 2 * int_of_float (ceil radius) + 1 in
     let pts = List.map (fun (x, y) -> (int_of_float (x +. 0.5), int_of_float (y +. 0.5))) (Polygon.points_on_path (float w /. 20.) (snd bs')) in
     ints.(0) <- k_brush; ints.(2) <- fl.n_points; ints.(3) <- List.length pts;
     fl.points <- pts :: fl.points; fl.n_points <- fl.n_points + List.length pts;
     floats.(6) <- opacity; floats.(7) <- radius);
  push fl (ints, floats, id)

(* scene list, then the reading-scene groups of its rewriting filters, then the (pages @ background) list *)
let flatten_scene (scene : Render.scene) (background : Render.scene) : u8 * int * i32 * i32 =
  let fl = new_flat () in
  List.iter (flatten_obj fl) scene;
  let reading = List.rev fl.reading in
  fl.reading <- [];
  let patches = ref [] in
  List.iter
    (fun (filter_rec, rewritten) ->
       let (ints, floats, _) = blank k_group_begin in
       ints.(26) <- 100;   (* COH_FILTER_READING_SCENE *)
       patches := (filter_rec, fl.n_recs) :: !patches;
       push fl (ints, floats, -1L);
       List.iter (flatten_obj fl ~cached:false) rewritten;
       push fl (blank k_group_end))
    reading;
  let before = fl.n_recs in
  List.iter (flatten_obj fl) background;
  fl.n_background <- fl.n_recs - before;
  let recs = Array.of_list (List.rev fl.recs) in
  List.iter (fun (filter_rec, group_rec) -> let (ints, _, _) = recs.(filter_rec) in ints.(22) <- group_rec) !patches;
  let sz = sizeof_object () in
  let objs = Array1.create int8_unsigned c_layout (sz * max 1 (Array.length recs)) in
  Array1.fill objs 0;
  Array.iteri (fun i (ints, floats, id) -> pack_object objs i ints floats id) recs;
  let objs = Array1.sub objs 0 (sz * Array.length recs) in
  let edges = i32_of_edges (List.concat (List.rev fl.edges)) in
  let pts = List.concat (List.rev fl.points) in
  let points = Array1.create int32 c_layout (2 * List.length pts) in
  List.iteri (fun i (x, y) -> points.{2 * i} <- Int32.of_int x; points.{2 * i + 1} <- Int32.of_int y) pts;
  (objs, fl.n_background, edges, points)

(* The rewriting filters need the scene BELOW them: [Rewritten f] receives [] above; callers that use affine / rgb /
   wireframe lenses register [Rewritten (fun _ -> rewritten_scene_below)] when they build the lens, which is where
   filters.ml builds that scene too (filters.ml:105-212: reading_scene closes over the transform). *)

(* ---- render.mli:211-217 ---- *)
let with_scene scene background f =
  let (objs, nbg, edges, points) = flatten_scene scene background in
  let h = scene_create (ctx ()) objs nbg edges points in
  match f h with
  | r -> scene_free (ctx ()) h; r
  | exception e -> scene_free (ctx ()) h; raise e

(* the sprite of the frame on `update` (the value render_frame returns; engine.ml:217-221 plots it) *)
let sprite_of_frame (update : shape_h) : Sprite.sprite =
  if update = 0L then Sprite.NullSprite
  else sprite_of_pixels (shape_export (ctx ()) update) (fb_read_sprite (ctx ()) update)

(* Render.drawable_of_rubberband is not exported by render.mli; the engine installs it here (one line) *)
let drawable_of_rubberband : (int -> int -> int -> int -> Render.scene) ref = ref (fun _ _ _ _ -> [])

let render_frame ?(display_selection = true) ?(topobjects = []) (_lmo : Id.idset) (view : Render.view) (update : Sprite.shape) : Sprite.sprite =
  ensure_canvas ();
  (* render.ml:1345-1365: rubberband @ selection @ topobjects @ view.scene, over pages @ background; the selection
     furniture is ordinary scene objects built by the reference's own Render.drawable_of_selection (render.mli:118) *)
  let selections = if display_selection then view.Render.selections else Render.null_selection in
  let rubberband = match view.Render.rubberband with None -> [] | Some (x0, y0, x1, y1) -> !drawable_of_rubberband x0 y0 x1 y1 in
  let front = rubberband @ Render.drawable_of_selection selections @ topobjects @ view.Render.scene in
  with_scene front (view.Render.pages @ view.Render.background)
    (fun h -> with_shape update (fun u -> render_frame_shape (ctx ()) h u 0; sprite_of_frame u))

let render_simple_scene (scene : Render.scene) (update : Sprite.shape) : Sprite.sprite =
  ensure_canvas ();
  with_scene scene [] (fun h -> with_shape update (fun u -> render_frame_shape (ctx ()) h u 0; sprite_of_frame u))

(* The same frame on every GPU of the box (engine start-up: [let gpus = multi_init [|0; 1; 2; 3|]]): the scene goes to
   every device, each renders its band of scanlines, the canvas bytes come from device 0 *)
let render_rect_rgb888_multi (gpus : multi) (view : Render.view) (x, y, w, h) (canvas_slice : u8) =
  let cw, ch = !canvas in
  multi_configure gpus cw ch;
  let (objs, nbg, edges, points) = flatten_scene view.Render.scene (view.Render.pages @ view.Render.background) in
  let s = multi_scene_create gpus objs nbg edges points in
  (try multi_render_frame gpus s (x, y, w, h) 0; multi_read_rgb888 gpus (x, y, w, h) canvas_slice
   with e -> multi_scene_free gpus s; raise e);
  multi_scene_free gpus s

(* engine.ml:208-221 render_rect: update = Sprite.box x y w h, result straight into the RGB888 canvas of wxgui.ml *)
let render_rect_rgb888 (view : Render.view) (x, y, w, h) (canvas_slice : u8) =
  ensure_canvas ();
  with_scene view.Render.scene (view.Render.pages @ view.Render.background)
    (fun s -> render_frame_rgb888 (ctx ()) s (x, y, w, h) canvas_slice)

(* ---- sprite.mli set algebra on Sprite.shape values (each call round-trips; keep handles for chains) ---- *)
let binop f a b = with_shape a (fun ha -> with_shape b (fun hb -> take_shape (f (ctx ()) ha hb)))
let ( ||| ) = binop shape_union
let ( --- ) = binop shape_difference
let ( &&& ) = binop shape_intersection
let bloat m n s = with_shape s (fun h -> take_shape (shape_bloat (ctx ()) h m n))
let erode m n s = with_shape s (fun h -> take_shape (shape_erode (ctx ()) h m n))
let translate_shape dx dy s = with_shape s (fun h -> take_shape (shape_translate (ctx ()) h dx dy))
let box x y w h = take_shape (shape_box (ctx ()) x y w h)

(* Render.plaindirty / alldirty (render.ml:1376-1391) *)
let dirty_region ~plain (so, mo, sn, mn) u =
  with_shape so (fun a -> with_shape mo (fun b -> with_shape sn (fun c -> with_shape mn (fun d -> with_shape u (fun hu ->
    take_shape (dirty_region_raw (ctx ()) (a, b, c, d) hu plain))))))

(* a Sprite.sprite as (device shape handle, pixels in span order); the handle is the caller's to free *)
let pixels_of_sprite (spr : Sprite.sprite) : i32 =
  let n = Sprite.sprite_card spr in
  let px = Array1.create int32 c_layout n in
  let i = ref 0 in
  Sprite.sprite_iter (fun _ _ c -> px.{!i} <- rgba8_of_colour c; incr i) spr;
  px

let fill_record (fill : Fill.fill) : u8 =
  let r = blank k_path in
  set_fill r (desc_of_fill fill);
  let (ints, floats, id) = r in
  let rec_ = Array1.create int8_unsigned c_layout (sizeof_object ()) in
  Array1.fill rec_ 0;
  pack_object rec_ 0 ints floats id;
  rec_

(* sprite.mli:96-125 *)
let shape_intersects a b = with_shape a (fun ha -> with_shape b (fun hb -> shape_intersects_raw (ctx ()) ha hb))
let portion (spr : Sprite.sprite) (shp : Sprite.shape) : Sprite.sprite =
  with_shape (Sprite.shape_of_sprite spr) (fun ha -> with_shape shp (fun hb ->
    let out = Array1.create int32 c_layout (shape_card (ctx ()) hb) in
    ignore (sprite_portion_raw (ctx ()) ha (pixels_of_sprite spr) hb out);
    sprite_of_pixels (shape_export (ctx ()) hb) out))
let fillshape (shp : Sprite.shape) (fill : Fill.fill) : Sprite.sprite =
  with_shape shp (fun h ->
    let out = Array1.create int32 c_layout (shape_card (ctx ()) h) in
    ignore (sprite_fillshape_raw (ctx ()) h (fill_record fill) out);
    sprite_of_pixels (shape_export (ctx ()) h) out)
type colour_map = Monochrome | Dissolve of int | Red_channel | Green_channel | Blue_channel   (* Sprite.sprite_map's closures, enumerated *)
let sprite_map (f : colour_map) (spr : Sprite.sprite) : Sprite.sprite =
  let px = pixels_of_sprite spr in
  let out = Array1.create int32 c_layout (Array1.dim px) in
  sprite_map_raw (ctx ()) (match f with Monochrome -> (0, 0) | Dissolve d -> (1, d) | Red_channel -> (2, 0) | Green_channel -> (3, 0) | Blue_channel -> (4, 0)) px out;
  sprite_of_pixels (flat_of_shape (Sprite.shape_of_sprite spr)) out

(* Polygon.polygon_sprite_edgelist fill shp edges winding (polygon.mli:58-59) *)
let polygon_sprite_edgelist (fill : Fill.fill) (shp : Sprite.shape) (edges : Polygon.edge list) winding : Sprite.sprite =
  with_shape shp (fun h ->
    let r = blank k_path in
    set_fill r (desc_of_fill fill);
    let (ints, floats, id) = r in
    let rec_ = Array1.create int8_unsigned c_layout (sizeof_object ()) in
    Array1.fill rec_ 0;
    pack_object rec_ 0 ints floats id;
    let n = shape_card (ctx ()) h in
    let out = Array1.create int32 c_layout n in
    ignore (polygon_sprite_raw (ctx ()) rec_ (i32_of_edges edges) (int_of_winding winding) h out);
    sprite_of_pixels (shape_export (ctx ()) h) out)

(* Convolve.convolve_sprite kernel sprite (convolve.mli:28-31) *)
let convolve_sprite (k : Convolve.kernel) (spr : Sprite.sprite) : Sprite.sprite =
  let shp = Sprite.shape_of_sprite spr in
  with_shape shp (fun h ->
    let n = shape_card (ctx ()) h in
    let px = Array1.create int32 c_layout n in
    let i = ref 0 in
    Sprite.sprite_iter (fun _ _ c -> px.{!i} <- rgba8_of_colour c; incr i) spr;   (* pixels in span order *)
    let r = Convolve.radius_of_kernel k in
    let grown = shape_bloat (ctx ()) h r r in
    let out = Array1.create int32 c_layout (shape_card (ctx ()) grown) in
    shape_free (ctx ()) grown;
    let rs = convolve_sprite_raw (ctx ()) (kind_of_kernel k, r) h px out in
    let s = sprite_of_pixels (shape_export (ctx ()) rs) out in
    shape_free (ctx ()) rs; s)

(* ---- N4 / N1: camlpy.mli through the library, and Wxgui.refresh_window without the canvas walk ---- *)
let wire_tuple = 0 and wire_unit = 1 and wire_int = 2 and wire_string = 3 and wire_bool = 4   (* camlpy.ml:26-30 *)

(* val marshall : marshallable -> string (camlpy.mli) *)
let marshall (m : Camlpy.marshallable) : string =
  let toks = ref [] and blob = Buffer.create 64 in
  let rec go = function
    | Camlpy.Unit -> toks := (wire_unit, 0L, 0L) :: !toks
    | Camlpy.Int i -> toks := (wire_int, Int64.of_int i, 0L) :: !toks
    | Camlpy.Bool b -> toks := (wire_bool, (if b then 1L else 0L), 0L) :: !toks
    | Camlpy.String st ->
        toks := (wire_string, Int64.of_int (String.length st), Int64.of_int (Buffer.length blob)) :: !toks;
        Buffer.add_string blob st
    | Camlpy.Tuple ls -> toks := (wire_tuple, Int64.of_int (List.length ls), 0L) :: !toks; List.iter go ls
  in
  go m;
  let toks = Array.of_list (List.rev !toks) in
  let n = Array.length toks in
  let kinds = Array1.create int32 c_layout n and values = Array1.create int64 c_layout n and offsets = Array1.create int64 c_layout n in
  Array.iteri (fun i (k, v, o) -> kinds.{i} <- Int32.of_int k; values.{i} <- v; offsets.{i} <- o) toks;
  host_wire_marshal kinds values offsets (Buffer.contents blob)

(* val unmarshall : string -> (int * marshallable) option (camlpy.mli).  Malformed data raises Invalid_data like
   camlpy.ml:84, 119-123 — a local exception there too (camlpy.mli does not export it), so this module has its own *)
exception Invalid_data
let unmarshall (str : string) : (int * Camlpy.marshallable) option =
  let cap = max 1 (String.length str) in   (* every token takes at least one byte *)
  let kinds = Array1.create int32 c_layout cap and values = Array1.create int64 c_layout cap and offsets = Array1.create int64 c_layout cap in
  let (taken, _) = try host_wire_unmarshal str kinds values offsets with Failure _ -> raise Invalid_data in
  if taken = 0 then None else begin
    let pos = ref 0 in
    let rec build () =
      let i = !pos in
      incr pos;
      let k = Int32.to_int kinds.{i} and v = Int64.to_int values.{i} in
      if k = wire_unit then Camlpy.Unit
      else if k = wire_int then Camlpy.Int v
      else if k = wire_bool then Camlpy.Bool (v <> 0)
      else if k = wire_string then Camlpy.String (String.sub str (Int64.to_int offsets.{i}) v)
      else begin
        let members = ref [] in
        for _k = 1 to v do members := build () :: !members done;
        Camlpy.Tuple (List.rev !members)
      end
    in
    Some (taken, build ())
  end

(* Wxgui.refresh_window window (xmin, ymin, xmax, ymax) (wxgui.ml:352-366) for a framebuffer kept on the GPU: the
   marshalled message, ready for Pytalk's send; "" for the rectangles the reference sends nothing for.  After
   Render.render_frame has run on the device (render_frame_box / scene_drag_object), this replaces plot_sprite +
   string_of_canvas_portion + Camlpy.marshall: no sprite is rebuilt, no canvas is walked. *)
let refresh_window_message (window : int) (rect : int * int * int * int) : string =
  wire_refresh_window (ctx ()) window rect
