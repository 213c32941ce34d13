(* coherence_gpu.ml — the OCaml side of the drop-in boundary (see INTEGRATION.md).
   NOT COMPILED HERE (no OCaml toolchain in this image); reviewed source.

   The reference's entry points keep their signatures; their bodies call these externals.
   Closures cannot cross the ABI: Fill.fill and the compositing operator are mapped to the
   descriptors of coh_object by [flatten_scene] (plain colour recovered as [fillsingle 0 0],
   as the reference itself does in sprite.ml:171; gradients / radials carry their parameters
   in a side table keyed by the fill record). *)
open Bigarray

type ctx = nativeint
type shape_h = int64   (* device-resident Sprite.shape; 0L = NullShape *)
type scene_h = int64
type i32 = (int32, int32_elt, c_layout) Array1.t
type u8 = (int, int8_unsigned_elt, c_layout) Array1.t

external init : int -> ctx = "coh_ml_init"
external shutdown : ctx -> unit = "coh_ml_shutdown"
external rgba8_of_colour : Colour.colour -> int32 = "coh_ml_rgba8_of_colour"
external colour_of_rgba8 : int32 -> Colour.colour = "coh_ml_colour_of_rgba8"
external shapeminshape : ctx -> i32 -> int -> shape_h * shape_h = "coh_ml_shapeminshape"
external shape_export : ctx -> shape_h -> i32 = "coh_ml_shape_export"
external shape_import : ctx -> i32 -> shape_h = "coh_ml_shape_import"
external shape_free : ctx -> shape_h -> unit = "coh_ml_shape_free"
external shape_union : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_union"
external shape_difference : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_difference"
external shape_intersection : ctx -> shape_h -> shape_h -> shape_h = "coh_ml_shape_intersection"
external shape_bloat : ctx -> shape_h -> int -> int -> shape_h = "coh_ml_shape_bloat"
external scene_create : ctx -> u8 -> int -> i32 -> i32 -> scene_h = "coh_ml_scene_create"
external scene_free : ctx -> scene_h -> unit = "coh_ml_scene_free"
external fb_configure : ctx -> int -> int -> int -> int -> unit = "coh_ml_fb_configure"
external render_frame_rgb888 : ctx -> scene_h -> int * int * int * int -> u8 -> unit = "coh_ml_render_frame_rgb888"
external cache_configure : ctx -> bool -> int64 -> unit = "coh_ml_cache_configure"
external cache_clear : ctx -> unit = "coh_ml_cache_clear"
external scene_drag_object : ctx -> scene_h -> int -> int -> int -> int * int * int * int = "coh_ml_scene_drag_object"
external scene_object_shape : ctx -> scene_h -> int -> shape_h * shape_h = "coh_ml_scene_object_shape"
external dirty_filter : ctx -> scene_h -> int -> shape_h -> shape_h = "coh_ml_dirty_filter"
external render_frame_shape : ctx -> scene_h -> shape_h -> unit = "coh_ml_render_frame_shape"
external read_rgb888 : ctx -> int * int * int * int -> u8 -> unit = "coh_ml_read_rgb888"
external convolve_sprite : ctx -> int * int -> shape_h -> (int32, int32_elt, c_layout) Array1.t -> (int32, int32_elt, c_layout) Array1.t -> shape_h = "coh_ml_convolve_sprite"

let the_ctx = lazy (init (-1))

(* Sprite.shape <-> flat records (y, nspans, (x, len) ...), rows grouped into vspans by consecutive y *)
let shape_of_flat (a : i32) : Sprite.shape =
  let n = Array1.dim a in
  if n = 0 then Sprite.NullShape else begin
    let rows = ref [] and i = ref 0 in
    while !i < n do
      let y = Int32.to_int a.{!i} and k = Int32.to_int a.{!i + 1} in
      let spans = List.init k (fun q -> (Int32.to_int a.{!i + 2 + 2 * q}, Int32.to_int a.{!i + 3 + 2 * q})) in
      rows := (y, spans) :: !rows;
      i := !i + 2 + 2 * k
    done;
    let vspans =
      List.fold_left
        (fun acc (y, spans) ->
           match acc with
           | (s, l, lines) :: rest when y = s - 1 -> (y, l + 1, spans :: lines) :: rest
           | _ -> (y, 1, [spans]) :: acc)
        [] !rows   (* rows are in decreasing y here, so vspans come out in increasing order *)
    in
    Sprite.boxshape (Sprite.Shape (Sprite.NoBounds, vspans))
  end

(* polygon.mli:55-59 *)
let shapeminshape_of_unsorted_edgelist (edges : Polygon.edge list) winding =
  let n = List.length edges in
  let a = Array1.create int32 c_layout (4 * n) in
  List.iteri
    (fun i (e : Polygon.edge) ->
       a.{4 * i} <- Int32.of_int e.Polygon.x0; a.{4 * i + 1} <- Int32.of_int e.Polygon.y0;
       a.{4 * i + 2} <- Int32.of_int e.Polygon.x1; a.{4 * i + 3} <- Int32.of_int e.Polygon.y1)
    edges;
  let ctx = Lazy.force the_ctx in
  let s, m = shapeminshape ctx a (match winding with Pdfgraphics.NonZero -> 0 | Pdfgraphics.EvenOdd -> 1) in
  let r = shape_of_flat (shape_export ctx s), shape_of_flat (shape_export ctx m) in
  shape_free ctx s; shape_free ctx m; r
