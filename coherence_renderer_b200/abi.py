"""ctypes binding of include/coherence_b200.h (the drop-in C ABI).

Loading fails loudly if the library has not been built; creating a context fails loudly
without a CUDA device.  Nothing in this module computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COH_LIB_PATH", os.path.join(_HERE, "libcoherence_b200.so"))  # override: kernel-variant experiments

COH_OBJ_PATH, COH_OBJ_PRIMITIVE, COH_OBJ_GROUP_BEGIN, COH_OBJ_GROUP_END, COH_OBJ_BRUSH, COH_OBJ_CPG = 0, 1, 2, 3, 4, 5
COH_OBJ_FILTER = 6
COH_BRUSH_GAUSSIAN, COH_BRUSH_DUMMY = 0, 1
COH_GEOM_PATH, COH_GEOM_NEXT = 0, 1
COH_FILTER_HOLE, COH_FILTER_MONOCHROME, COH_FILTER_BLUR, COH_FILTER_SCENE, COH_FILTER_MINUS, COH_FILTER_SMEAR, COH_FILTER_READING_SCENE = 1, 2, 3, 4, 5, 6, 100
COH_CPG_UNION, COH_CPG_INTERSECTION, COH_CPG_SUBTRACTION, COH_CPG_EXCLUSIVEOR = 0, 1, 2, 3
COH_NONZERO, COH_EVENODD = 0, 1
COH_FILL_PLAIN, COH_FILL_AXIAL, COH_FILL_RADIAL = 0, 1, 2
COH_FILL_EXT_S, COH_FILL_EXT_E = 1, 2
COH_RENDER_RECORD_U = 1
COH_CONV_UNIT, COH_CONV_GAUSSIAN = 1, 2


class CohObject(C.Structure):
    """struct coh_object (include/coherence_b200.h)."""

    _fields_ = [
        ("kind", C.c_int32), ("winding", C.c_int32), ("first", C.c_int32), ("count", C.c_int32),
        ("fill_kind", C.c_int32), ("colour0", C.c_uint32), ("colour1", C.c_uint32), ("fill_flags", C.c_int32),
        ("pretrans", C.c_int32), ("dx", C.c_int32), ("dy", C.c_int32), ("bounds", C.c_int32 * 4),
        ("prim", C.c_int32 * 4), ("prim_null", C.c_int32), ("convolve", C.c_int32), ("sprite_winding", C.c_int32), ("id", C.c_int64),
        ("fparam", C.c_double * 6), ("brush_opacity", C.c_double), ("brush_radius", C.c_double),
        ("first2", C.c_int32), ("count2", C.c_int32), ("winding2", C.c_int32), ("cpg_op", C.c_int32),
        ("filter_kind", C.c_int32), ("filter_kernel", C.c_int32),
    ]


# every symbol include/coherence_b200.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "coh_init", "coh_shutdown", "coh_last_error", "coh_device_name", "coh_stream", "coh_launch_count",
    "coh_set_stream", "coh_set_timing", "coh_get_timing", "coh_fb_attach", "coh_set_option",
    "coh_colour_of_rgba8", "coh_rgba8_of_colour", "coh_shapeminshape_of_edgelist", "coh_edgelist_of_path", "coh_shapeminshape_of_path", "coh_brush_shape", "coh_brush_sprite", "coh_brush_smear", "coh_polygon_opacity",
    "coh_polygon_sprite", "coh_shape_box", "coh_shape_import", "coh_shape_export_size", "coh_shape_export",
    "coh_shape_bounds", "coh_shape_card", "coh_shape_free", "coh_shape_union", "coh_shape_difference",
    "coh_shape_intersection", "coh_shape_translate", "coh_shape_bloat", "coh_shape_erode", "coh_scene_create",
    "coh_scene_free", "coh_fb_configure", "coh_render_frame", "coh_render_frame_shape", "coh_scene_translate_object", "coh_render_uncovered", "coh_sync",
    "coh_fb_device_ptr", "coh_fb_read_rgba", "coh_fb_read_rgb888", "coh_fb_read_sprite", "coh_fb_read_rgba_async", "coh_fb_read_wait", "coh_fb_set_peers", "coh_mem_in_use",
    "coh_host_edgelist_of_subpath", "coh_host_brush_points", "coh_host_smear_points",
    "coh_host_strokepath", "coh_host_bounds_stroke", "coh_strokepath", "coh_shapeminshape_of_stroke",
    "coh_cache_configure", "coh_cache_clear", "coh_cache_stats", "coh_cache_sprite_stats", "coh_cache_addshape", "coh_cache_getshape",
    "coh_cache_addtranslation", "coh_dirty_region", "coh_scene_drag_object", "coh_dirty_filter", "coh_scene_object_shape", "coh_convolve_sprite",
    "coh_multi_init", "coh_multi_shutdown", "coh_multi_last_error", "coh_multi_device_count", "coh_multi_ctx", "coh_multi_configure",
    "coh_multi_scene_create", "coh_multi_scene_free", "coh_multi_scene_translate_object", "coh_multi_render_frame", "coh_multi_sync",
    "coh_multi_fb_read_rgba", "coh_multi_fb_read_rgb888", "coh_fb_alloc_shared", "coh_fb_open_peer", "coh_frame_signal", "coh_frame_wait",
    "coh_shape_intersects", "coh_sprite_portion", "coh_sprite_fillshape", "coh_sprite_map", "coh_sprite_map_coords_fill",
    "coh_host_wire_marshal", "coh_host_wire_unmarshal", "coh_host_wire_refresh_window", "coh_wire_refresh_window",
]

_lib = None


class CohError(RuntimeError):
    """The OCaml stub raises `Failure msg` for these (reference convention: failwith)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CohError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        L.coh_last_error.restype = C.c_char_p
        L.coh_last_error.argtypes = [C.c_void_p]
        L.coh_stream.restype = C.c_void_p
        L.coh_stream.argtypes = [C.c_void_p]
        L.coh_fb_device_ptr.restype = C.c_void_p
        L.coh_fb_device_ptr.argtypes = [C.c_void_p]
        L.coh_launch_count.restype = C.c_int64
        L.coh_launch_count.argtypes = [C.c_void_p]
        L.coh_rgba8_of_colour.restype = C.c_uint32
        L.coh_colour_of_rgba8.restype = C.c_int32
        L.coh_host_edgelist_of_subpath.restype = C.c_int64
        L.coh_host_brush_points.restype = C.c_int64
        L.coh_host_smear_points.restype = C.c_int64
        L.coh_host_strokepath.restype = C.c_int64
        L.coh_host_wire_marshal.restype = C.c_int64
        L.coh_host_wire_refresh_window.restype = C.c_int64
        L.coh_multi_last_error.restype = C.c_char_p
        L.coh_multi_last_error.argtypes = [C.c_void_p]
        L.coh_multi_ctx.restype = C.c_void_p
        L.coh_multi_ctx.argtypes = [C.c_void_p, C.c_int32]
        _lib = L
    return _lib


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Context:
    """One GPU, one scanline band (coh_init / coh_shutdown)."""

    def __init__(self, device=-1, _borrowed=None):
        self._h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)   # a device's context of a MultiContext
            return
        rc = lib().coh_init(device, C.byref(self._h))
        if rc != 0:
            raise CohError(lib().coh_last_error(None).decode())
        for kv in filter(None, os.environ.get("COH_OPTIONS", "").split(",")):   # A/B measurements: COH_OPTIONS=pdl=0,ab=1
            name, _, value = kv.partition("=")
            self.set_option(name.strip(), int(value))

    def close(self):
        if self._h and self._owned:
            lib().coh_shutdown(self._h)
        self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise CohError(lib().coh_last_error(self._h).decode())

    # -- info
    def device_name(self):
        b = C.create_string_buffer(256)
        self._chk(lib().coh_device_name(self._h, b, 256))
        return b.value.decode()

    def stream(self):
        return lib().coh_stream(self._h)

    def launch_count(self):
        return lib().coh_launch_count(self._h)

    def set_stream(self, cuda_stream):
        self._chk(lib().coh_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_timing(self, on):
        self._chk(lib().coh_set_timing(self._h, 1 if on else 0))

    def get_timing(self):
        w, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._chk(lib().coh_get_timing(self._h, C.byref(w), C.byref(b), C.byref(n)))
        return w.value, b.value, n.value

    def set_option(self, name, value):
        self._chk(lib().coh_set_option(self._h, name.encode(), int(value)))

    def fb_attach(self, device_ptr):
        self._chk(lib().coh_fb_attach(self._h, C.c_void_p(device_ptr)))

    def frame_signal(self, target_fbs, slot, epoch):
        """Counter `slot` of every target framebuffer (fb_open_peer pointers / own fb_device_ptr) becomes `epoch` behind the
        work issued so far on the stream."""
        arr = (C.c_void_p * max(len(target_fbs), 1))(*[C.c_void_p(int(p)) for p in target_fbs])
        self._chk(lib().coh_frame_signal(self._h, len(target_fbs), arr, int(slot), int(epoch)))

    def frame_wait(self, slots, epoch):
        """The stream waits on the device until the listed counters of the own shared framebuffer have reached `epoch`."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        self._chk(lib().coh_frame_wait(self._h, len(sl), _i32p(sl), int(epoch)))

    def mem_in_use(self):
        n = C.c_int64()
        self._chk(lib().coh_mem_in_use(self._h, C.byref(n)))
        return n.value

    def sync(self):
        self._chk(lib().coh_sync(self._h))

    # -- shapes
    def shape_import(self, flat):
        a = np.ascontiguousarray(flat, dtype=np.int32)
        h = C.c_uint64()
        self._chk(lib().coh_shape_import(self._h, _i32p(a), C.c_int64(len(a)), C.byref(h)))
        return h.value

    def shape_export(self, h):
        n = C.c_int64()
        self._chk(lib().coh_shape_export_size(self._h, C.c_uint64(h), C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.int32)
        m = C.c_int64()
        self._chk(lib().coh_shape_export(self._h, C.c_uint64(h), _i32p(out), C.c_int64(n.value), C.byref(m)))
        return out[: m.value]

    def shape_free(self, h):
        if h:
            self._chk(lib().coh_shape_free(self._h, C.c_uint64(h)))

    def shape_box(self, x, y, w, h):
        o = C.c_uint64()
        self._chk(lib().coh_shape_box(self._h, x, y, w, h, C.byref(o)))
        return o.value

    def shape_bounds(self, h):
        box = (C.c_int32 * 4)()
        isnull = C.c_int32()
        self._chk(lib().coh_shape_bounds(self._h, C.c_uint64(h), box, C.byref(isnull)))
        return None if isnull.value else tuple(box)

    def shape_card(self, h):
        n = C.c_int64()
        self._chk(lib().coh_shape_card(self._h, C.c_uint64(h), C.byref(n)))
        return n.value

    def _binop(self, fn, a, b):
        o = C.c_uint64()
        self._chk(fn(self._h, C.c_uint64(a), C.c_uint64(b), C.byref(o)))
        return o.value

    def shape_union(self, a, b):
        return self._binop(lib().coh_shape_union, a, b)

    def shape_difference(self, a, b):
        return self._binop(lib().coh_shape_difference, a, b)

    def shape_intersection(self, a, b):
        return self._binop(lib().coh_shape_intersection, a, b)

    def _unop(self, fn, a, m, n):
        o = C.c_uint64()
        self._chk(fn(self._h, C.c_uint64(a), m, n, C.byref(o)))
        return o.value

    def shape_translate(self, a, dx, dy):
        return self._unop(lib().coh_shape_translate, a, dx, dy)

    def shape_bloat(self, a, m, n):
        return self._unop(lib().coh_shape_bloat, a, m, n)

    def shape_erode(self, a, m, n):
        return self._unop(lib().coh_shape_erode, a, m, n)

    # -- polygon
    def shapeminshape_of_edgelist(self, edges, winding):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
        s, m = C.c_uint64(), C.c_uint64()
        self._chk(lib().coh_shapeminshape_of_edgelist(self._h, _i32p(e), len(e), winding, C.byref(s), C.byref(m)))
        return s.value, m.value

    def edgelist_of_path(self, segs):
        """Polygon.edgelist_of_path on the device for the segments of a path (N2)."""
        rec = _seg_records(segs)
        cap = 64 * len(rec) + 64
        while True:
            out = np.zeros((cap, 4), dtype=np.int32)
            n = C.c_int64()
            self._chk(lib().coh_edgelist_of_path(self._h, rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), _i32p(out), C.c_int64(cap), C.byref(n)))
            if n.value <= cap:
                return out[:n.value]
            cap = int(n.value)

    def shapeminshape_of_path(self, segs, winding):
        """Polygon.shapeminshape_polygon: flattening and scan conversion both on the device."""
        rec = _seg_records(segs)
        s, m = C.c_uint64(), C.c_uint64()
        self._chk(lib().coh_shapeminshape_of_path(self._h, rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), winding, C.byref(s), C.byref(m)))
        return s.value, m.value

    def strokepath(self, spec, subpaths):
        """Shapes.strokepath: the stroke's outline (host stroker) flattened on the device, edges sorted by maximum y.
        Returns (edges, winding rule of the outline)."""
        rec, cnt = _path_records(subpaths)
        cap = 256 * len(rec) + 256
        while True:
            out = np.zeros((cap, 4), dtype=np.int32)
            n, w = C.c_int64(), C.c_int32()
            self._chk(lib().coh_strokepath(self._h, C.byref(spec), rec.ctypes.data_as(C.POINTER(C.c_double)), _i32p(cnt), len(cnt),
                                           _i32p(out), C.c_int64(cap), C.byref(n), C.byref(w)))
            if n.value <= cap:
                return out[:n.value], w.value
            cap = int(n.value)

    def shapeminshape_of_stroke(self, spec, subpaths):
        """Shape and minshape of a stroked path: stroker on the host, flattening and scan conversion on the device."""
        rec, cnt = _path_records(subpaths)
        s, m = C.c_uint64(), C.c_uint64()
        self._chk(lib().coh_shapeminshape_of_stroke(self._h, C.byref(spec), rec.ctypes.data_as(C.POINTER(C.c_double)), _i32p(cnt), len(cnt), C.byref(s), C.byref(m)))
        return s.value, m.value

    def polygon_opacity(self, edges, winding, shp):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
        cap = self.shape_card(shp)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = C.c_int64()
        self._chk(lib().coh_polygon_opacity(self._h, _i32p(e), len(e), winding, C.c_uint64(shp), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def polygon_sprite(self, fill_obj, edges, winding, shp):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
        cap = self.shape_card(shp)
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        n = C.c_int64()
        self._chk(lib().coh_polygon_sprite(self._h, C.byref(fill_obj), _i32p(e), len(e), winding, C.c_uint64(shp), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    # -- brush strokes outside a scene (brush.mli:20-27): brush_obj = a BRUSH CohObject, points = rounded stamp points
    def brush_shape(self, brush_obj, points):
        p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
        h = C.c_uint64()
        self._chk(lib().coh_brush_shape(self._h, C.byref(brush_obj), _i32p(p), len(p), C.byref(h)))
        return h.value

    def brush_sprite(self, brush_obj, points, shp):
        p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
        cap = self.shape_card(shp)
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        n = C.c_int64()
        self._chk(lib().coh_brush_sprite(self._h, C.byref(brush_obj), _i32p(p), len(p), C.c_uint64(shp), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def brush_smear(self, shape, rgba, brush_obj, points, smear_points):
        """Brush.smear: returns (shape handle of the result, rgba per pixel)."""
        p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
        q = np.ascontiguousarray(smear_points, dtype=np.int32).reshape(-1, 2)
        src = np.ascontiguousarray(rgba, dtype=np.uint32)
        bs = self.brush_shape(brush_obj, p)
        un = self.shape_union(shape, bs)
        cap = self.shape_card(un)
        self.shape_free(bs)
        self.shape_free(un)
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        o, n = C.c_uint64(), C.c_int64()
        self._chk(lib().coh_brush_smear(self._h, C.c_uint64(shape), src.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(brush_obj), _i32p(p), len(p), _i32p(q), len(q),
                                        C.byref(o), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return o.value, out[: n.value]

    # -- sprites: (shape handle, RGBA8 per pixel in span order)
    def shape_intersects(self, a, b):
        y = C.c_int32()
        self._chk(lib().coh_shape_intersects(self._h, C.c_uint64(a), C.c_uint64(b), C.byref(y)))
        return bool(y.value)

    def _sprite_out(self, shape):
        cap = self.shape_card(shape) if shape else 0
        return cap, np.zeros(max(cap, 1), dtype=np.uint32), C.c_int64()

    def sprite_portion(self, shape, rgba, sub):
        src = np.ascontiguousarray(rgba, dtype=np.uint32)
        cap, out, n = self._sprite_out(sub)
        self._chk(lib().coh_sprite_portion(self._h, C.c_uint64(shape), src.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(sub), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def sprite_fillshape(self, shape, fill_obj):
        cap, out, n = self._sprite_out(shape)
        self._chk(lib().coh_sprite_fillshape(self._h, C.c_uint64(shape), C.byref(fill_obj), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def sprite_map(self, op, rgba, arg=0):
        src = np.ascontiguousarray(rgba, dtype=np.uint32)
        out = np.zeros(max(len(src), 1), dtype=np.uint32)
        code = {"monochrome": 0, "dissolve": 1, "red_channel": 2, "green_channel": 3, "blue_channel": 4}[op]
        self._chk(lib().coh_sprite_map(self._h, code, arg, src.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(len(src)), out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out[: len(src)]

    def sprite_map_coords_fill(self, shape, fill_obj, rgba):
        src = np.ascontiguousarray(rgba, dtype=np.uint32)
        cap, out, n = self._sprite_out(shape)
        self._chk(lib().coh_sprite_map_coords_fill(self._h, C.c_uint64(shape), C.byref(fill_obj), src.ctypes.data_as(C.POINTER(C.c_uint32)), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def convolve_sprite(self, kernel, r, shape, rgba):
        """Convolve.convolve_sprite (mkunit r | mkgaussian r): returns (shape handle of the result, rgba per pixel)."""
        kind = {"unit": COH_CONV_UNIT, "gaussian": COH_CONV_GAUSSIAN}[kernel]
        src = np.ascontiguousarray(rgba, dtype=np.uint32)
        bl = self.shape_bloat(shape, r, r)
        cap = self.shape_card(bl)
        self.shape_free(bl)
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        o, n = C.c_uint64(), C.c_int64()
        self._chk(lib().coh_convolve_sprite(self._h, kind, r, C.c_uint64(shape), src.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(o), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return o.value, out[: n.value]

    # -- scenes / rendering
    def scene_create(self, objs, n_background, edges, points):
        """objs: ctypes array of CohObject (scene objects then n_background background objects)."""
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
        p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
        h = C.c_uint64()
        self._chk(lib().coh_scene_create(self._h, objs, len(objs), n_background, _i32p(e), len(e), _i32p(p), len(p), C.byref(h)))
        return h.value

    def scene_free(self, h):
        if h:
            self._chk(lib().coh_scene_free(self._h, C.c_uint64(h)))

    def fb_configure(self, width, height, band_y0=0, band_y1=None):
        self._chk(lib().coh_fb_configure(self._h, width, height, band_y0, height if band_y1 is None else band_y1))
        self._wh = (width, height)

    def render_frame(self, scene, update, flags=0):
        ux, uy, uw, uh = update
        self._chk(lib().coh_render_frame(self._h, C.c_uint64(scene), ux, uy, uw, uh, flags))

    def render_frame_shape(self, scene, update_shape, flags=0):
        self._chk(lib().coh_render_frame_shape(self._h, C.c_uint64(scene), C.c_uint64(update_shape), flags))

    def scene_translate_object(self, scene, obj_index, dx, dy):
        self._chk(lib().coh_scene_translate_object(self._h, C.c_uint64(scene), obj_index, dx, dy))

    def scene_object_shape(self, scene, obj_index):
        s, m = C.c_uint64(), C.c_uint64()
        self._chk(lib().coh_scene_object_shape(self._h, C.c_uint64(scene), obj_index, C.byref(s), C.byref(m)))
        return s.value, m.value

    def scene_drag_object(self, scene, obj_index, dx, dy, flags=0):
        """One drag step (translate, dirty region from cached span sets, render it); returns the dirty pixel box."""
        bb = (C.c_int32 * 4)()
        self._chk(lib().coh_scene_drag_object(self._h, C.c_uint64(scene), obj_index, dx, dy, flags, bb))
        return tuple(bb)

    def dirty_filter(self, scene, lmo_index, initial_dirty):
        o = C.c_uint64()
        self._chk(lib().coh_dirty_filter(self._h, C.c_uint64(scene), lmo_index, C.c_uint64(initial_dirty), C.byref(o)))
        return o.value

    def dirty_region(self, shp_o, min_o, shp_n, min_n, u, plain):
        o = C.c_uint64()
        self._chk(lib().coh_dirty_region(self._h, C.c_uint64(shp_o), C.c_uint64(min_o), C.c_uint64(shp_n), C.c_uint64(min_n), C.c_uint64(u), 1 if plain else 0, C.byref(o)))
        return o.value

    def cache_configure(self, usecache=True, max_bytes=0):
        self._chk(lib().coh_cache_configure(self._h, 1 if usecache else 0, C.c_int64(max_bytes)))

    def cache_clear(self):
        self._chk(lib().coh_cache_clear(self._h))

    def cache_stats(self):
        out = (C.c_int64 * 4)()
        self._chk(lib().coh_cache_stats(self._h, out))
        return {"shape_hits": out[0], "shape_misses": out[1], "bytes": out[2], "entries": out[3]}

    def cache_sprite_stats(self, scene):
        out = (C.c_int64 * 4)()
        self._chk(lib().coh_cache_sprite_stats(self._h, C.c_uint64(scene), out))
        return {"sprite_hits": out[0], "sprite_fills": out[1], "bytes": out[2], "entries": out[3]}

    def cache_addshape(self, oid, shape, minshape):
        self._chk(lib().coh_cache_addshape(self._h, C.c_int64(oid), C.c_uint64(shape), C.c_uint64(minshape)))

    def cache_getshape(self, oid):
        s, m, f = C.c_uint64(), C.c_uint64(), C.c_int32()
        self._chk(lib().coh_cache_getshape(self._h, C.c_int64(oid), C.byref(s), C.byref(m), C.byref(f)))
        return (s.value, m.value) if f.value else None

    def cache_addtranslation(self, oid, target, dx, dy):
        self._chk(lib().coh_cache_addtranslation(self._h, C.c_int64(oid), C.c_int64(target), dx, dy))

    def render_uncovered(self):
        o = C.c_uint64()
        self._chk(lib().coh_render_uncovered(self._h, C.byref(o)))
        return o.value

    def fb_device_ptr(self):
        return lib().coh_fb_device_ptr(self._h)

    def fb_read_rgba(self, x, y, w, h, out=None):
        if out is None:
            out = np.zeros((h, w), dtype=np.uint32)
        self._chk(lib().coh_fb_read_rgba(self._h, x, y, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def fb_set_peers(self, peer_ptrs):
        """peer_ptrs: device pointers (ints) of the other ranks' framebuffers, peer-mapped into this process."""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(p) for p in peer_ptrs])
        self._chk(lib().coh_fb_set_peers(self._h, len(peer_ptrs), arr))

    def fb_alloc_shared(self):
        """A framebuffer that other processes can map: returns its 64-byte CUDA IPC handle."""
        h = (C.c_uint8 * 64)()
        self._chk(lib().coh_fb_alloc_shared(self._h, h))
        return bytes(h)

    def fb_open_peer(self, handle):
        """Map another process's shared framebuffer; returns the device pointer for fb_set_peers."""
        buf = (C.c_uint8 * 64)(*handle)
        p = C.c_void_p()
        self._chk(lib().coh_fb_open_peer(self._h, buf, C.byref(p)))
        return p.value

    def fb_read_rgba_async(self, x, y, w, h, out):
        """out: pinned host array of h*w uint32; valid after fb_read_wait()."""
        self._chk(lib().coh_fb_read_rgba_async(self._h, x, y, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))

    def fb_read_wait(self):
        self._chk(lib().coh_fb_read_wait(self._h))

    def fb_read_sprite(self, update_shape):
        """The frame's pixels on `update_shape`, RGBA8 per pixel in canonical span order (Render.render_frame's sprite)."""
        cap = self.shape_card(update_shape) if update_shape else 0
        out = np.zeros(max(cap, 1), dtype=np.uint32)
        n = C.c_int64()
        self._chk(lib().coh_fb_read_sprite(self._h, C.c_uint64(update_shape), out.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(cap), C.byref(n)))
        return out[: n.value]

    def fb_read_rgb888(self, x, y, w, h):
        out = np.zeros((h, w, 3), dtype=np.uint8)
        self._chk(lib().coh_fb_read_rgb888(self._h, x, y, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def wire_refresh_window(self, window, xmin, ymin, xmax, ymax):
        """Wxgui.refresh_window's marshalled message (wxgui.ml:352-366) with the pixels of the framebuffer; b"" where the
        reference sends nothing."""
        n = C.c_int64()
        self._chk(lib().coh_wire_refresh_window(self._h, window, xmin, ymin, xmax, ymax, None, C.c_int64(0), C.byref(n)))
        if n.value == 0:
            return b""
        out = np.zeros(n.value, dtype=np.uint8)
        self._chk(lib().coh_wire_refresh_window(self._h, window, xmin, ymin, xmax, ymax, out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(n.value), C.byref(n)))
        return out.tobytes()


class MultiContext:
    """Several GPUs of one box from one process (coh_multi_*): every device renders its band of scanlines and stores
    its pixels into every framebuffer over NVLink."""

    def __init__(self, n_devices, device_ids=None):
        self._h = C.c_void_p()
        ids = (C.c_int32 * n_devices)(*device_ids) if device_ids else None
        if lib().coh_multi_init(n_devices, ids, C.byref(self._h)) != 0:
            raise CohError(lib().coh_multi_last_error(None).decode())
        self.n = n_devices

    def _chk(self, rc):
        if rc != 0:
            raise CohError(lib().coh_multi_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().coh_multi_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def ctx(self, i):
        return Context(_borrowed=lib().coh_multi_ctx(self._h, i))

    def configure(self, width, height, cuts=None):
        arr = (C.c_int32 * (self.n + 1))(*cuts) if cuts is not None else None
        self._chk(lib().coh_multi_configure(self._h, width, height, arr))

    def scene_create(self, objs, n_background, edges, points):
        e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
        p = np.ascontiguousarray(points, dtype=np.int32).reshape(-1, 2)
        h = C.c_uint64()
        self._chk(lib().coh_multi_scene_create(self._h, objs, len(objs), n_background, _i32p(e), len(e), _i32p(p), len(p), C.byref(h)))
        return h.value

    def scene_free(self, h):
        if h:
            self._chk(lib().coh_multi_scene_free(self._h, C.c_uint64(h)))

    def scene_translate_object(self, scene, obj_index, dx, dy):
        self._chk(lib().coh_multi_scene_translate_object(self._h, C.c_uint64(scene), obj_index, dx, dy))

    def render_frame(self, scene, update, flags=0):
        ux, uy, uw, uh = update
        self._chk(lib().coh_multi_render_frame(self._h, C.c_uint64(scene), ux, uy, uw, uh, flags))

    def sync(self):
        self._chk(lib().coh_multi_sync(self._h))

    def fb_read_rgba(self, x, y, w, h, out=None):
        if out is None:
            out = np.zeros((h, w), dtype=np.uint32)
        self._chk(lib().coh_multi_fb_read_rgba(self._h, x, y, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def fb_read_rgb888(self, x, y, w, h):
        out = np.zeros((h, w, 3), dtype=np.uint8)
        self._chk(lib().coh_multi_fb_read_rgb888(self._h, x, y, w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out


def _seg_records(segs):
    rec = np.zeros((len(segs), 9), dtype=np.float64)
    for i, s in enumerate(segs):
        rec[i, 0] = 0.0 if s[0] == "L" else 1.0
        for k, p in enumerate(s[1:]):
            rec[i, 1 + 2 * k], rec[i, 2 + 2 * k] = p
    return rec


def _path_records(subpaths):
    """Segment records of all the subpaths of a path in order + the number of segments of each."""
    cnt = np.array([len(sp) for sp in subpaths], dtype=np.int32)
    rec = _seg_records([s for sp in subpaths for s in sp])
    return rec, cnt


CAP_BUTT, CAP_ROUND, CAP_PROJECTING = 0, 1, 2
JOIN_ROUND, JOIN_MITRED, JOIN_BEVEL = 0, 1, 2


class StrokeSpec(C.Structure):
    """coh_strokespec (shapes.ml:166-171)."""
    _fields_ = [("startcap", C.c_int32), ("join", C.c_int32), ("endcap", C.c_int32), ("reserved", C.c_int32),
                ("mitrelimit", C.c_double), ("linewidth", C.c_double)]


def strokespec(startcap, join, endcap, mitrelimit, linewidth):
    return StrokeSpec(startcap, join, endcap, 0, mitrelimit, linewidth)


def host_strokepath(spec, subpaths):
    """Shapes.strokepath_polygon through the library's host stroker: (outline records (n, 9), segments per outline
    subpath, winding rule)."""
    rec, cnt = _path_records(subpaths)
    m, w = C.c_int32(), C.c_int32()
    counts = np.zeros(max(len(cnt), 1), dtype=np.int32)
    n = lib().coh_host_strokepath(C.byref(spec), rec.ctypes.data_as(C.POINTER(C.c_double)), _i32p(cnt), len(cnt),
                                  None, C.c_int64(0), _i32p(counts), 0, C.byref(m), C.byref(w))
    if n < 0:
        raise CohError("Shapes.joinsegments: a rail that ends in a curve cannot be joined")
    out = np.zeros((max(int(n), 1), 9), dtype=np.float64)
    lib().coh_host_strokepath(C.byref(spec), rec.ctypes.data_as(C.POINTER(C.c_double)), _i32p(cnt), len(cnt),
                              out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(n), _i32p(counts), len(counts), C.byref(m), C.byref(w))
    return out[:n], counts[: m.value], w.value


def host_bounds_stroke(spec, subpaths):
    """Shapes.bounds_stroke: (xmin, xmax, ymin, ymax) of the stroke in pixels."""
    rec, cnt = _path_records(subpaths)
    out = np.zeros(4, dtype=np.int32)
    if lib().coh_host_bounds_stroke(C.byref(spec), rec.ctypes.data_as(C.POINTER(C.c_double)), _i32p(cnt), len(cnt), _i32p(out)) != 0:
        raise CohError("Polygon2.bounds_polygon: Malformed (empty) path")
    return tuple(int(v) for v in out)


# ---- the front end's socket format (camlpy.mli): Python values <-> Camlpy.marshallable as pycaml.py maps them
# (None = Unit, bool = Bool, int = Int, bytes / str = String, list / tuple = Tuple)
WIRE_TUPLE, WIRE_UNIT, WIRE_INT, WIRE_STRING, WIRE_BOOL = 0, 1, 2, 3, 4


def _wire_tokens(m, kinds, values, offsets, blob):
    if m is None:
        kinds.append(WIRE_UNIT), values.append(0), offsets.append(0)
    elif isinstance(m, bool):
        kinds.append(WIRE_BOOL), values.append(int(m)), offsets.append(0)
    elif isinstance(m, int):
        kinds.append(WIRE_INT), values.append(m), offsets.append(0)
    elif isinstance(m, (bytes, bytearray, str)):
        b = m.encode("latin-1") if isinstance(m, str) else bytes(m)
        kinds.append(WIRE_STRING), values.append(len(b)), offsets.append(len(blob))
        blob.extend(b)
    elif isinstance(m, (list, tuple)):
        kinds.append(WIRE_TUPLE), values.append(len(m)), offsets.append(0)
        for e in m:
            _wire_tokens(e, kinds, values, offsets, blob)
    else:
        raise CohError("Invalid Data")


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def host_wire_marshal_tokens(kinds, values, offsets, blob=b""):
    """coh_host_wire_marshal on a raw token list (None when the list is not exactly one value)."""
    k, v, o = np.asarray(kinds, dtype=np.int32), np.asarray(values, dtype=np.int64), np.asarray(offsets, dtype=np.int64)
    sb = np.frombuffer(bytes(blob) + b"\0", dtype=np.uint8)
    n = lib().coh_host_wire_marshal(_i32p(k), _i64p(v), _i64p(o), len(k), sb.ctypes.data_as(C.POINTER(C.c_uint8)), None, C.c_int64(0))
    if n < 0:
        return None
    out = np.zeros(n, dtype=np.uint8)
    lib().coh_host_wire_marshal(_i32p(k), _i64p(v), _i64p(o), len(k), sb.ctypes.data_as(C.POINTER(C.c_uint8)), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int64(n))
    return out.tobytes()


def host_wire_marshal(m):
    """Camlpy.marshall (camlpy.ml:77-82) of a Python value."""
    kinds, values, offsets, blob = [], [], [], bytearray()
    _wire_tokens(m, kinds, values, offsets, blob)
    return host_wire_marshal_tokens(kinds, values, offsets, blob)


def host_wire_unmarshal(data):
    """Camlpy.unmarshall (camlpy.ml:106-124): None while the message is incomplete, else (bytes taken, value); CohError for
    Invalid_data.  Strings come back as bytes."""
    buf = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8)
    nt, taken = C.c_int32(), C.c_int64()
    bp = buf.ctypes.data_as(C.POINTER(C.c_uint8))
    if lib().coh_host_wire_unmarshal(bp, C.c_int64(len(data)), None, None, None, 0, C.byref(nt), C.byref(taken)) != 0:
        raise CohError("Invalid_data")
    if taken.value == 0:
        return None
    k, v, o = np.zeros(nt.value, dtype=np.int32), np.zeros(nt.value, dtype=np.int64), np.zeros(nt.value, dtype=np.int64)
    lib().coh_host_wire_unmarshal(bp, C.c_int64(len(data)), _i32p(k), _i64p(v), _i64p(o), nt.value, C.byref(nt), C.byref(taken))
    pos = [0]

    def build():
        i = pos[0]
        pos[0] += 1
        if k[i] == WIRE_UNIT:
            return None
        if k[i] == WIRE_BOOL:
            return bool(v[i])
        if k[i] == WIRE_INT:
            return int(v[i])
        if k[i] == WIRE_STRING:
            return bytes(data[int(o[i]) : int(o[i]) + int(v[i])])
        return [build() for _ in range(int(v[i]))]

    return taken.value, build()


def host_wire_refresh_window(window, xmin, ymin, xmax, ymax):
    """(size of the whole RefreshWindow message, the bytes in front of its pixels); (0, b"") where nothing is sent."""
    hdr = np.zeros(64, dtype=np.uint8)
    hl = C.c_int32()
    n = lib().coh_host_wire_refresh_window(window, xmin, ymin, xmax, ymax, hdr.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(hl))
    if n < 0:
        raise CohError("refresh_window: not a rectangle (wxgui.ml:335)")
    return int(n), hdr[: hl.value].tobytes()


def host_edgelist_of_subpath(segs):
    """Polygon.edgelist_of_path for one subpath through the library's host-side geometry."""
    rec = _seg_records(segs)
    cap = 64 * len(segs) + 64
    while True:
        out = np.zeros((cap, 4), dtype=np.int32)
        n = lib().coh_host_edgelist_of_subpath(rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), _i32p(out), C.c_int64(cap))
        if n <= cap:
            return out[:n]
        cap = int(n)


def host_smear_points(segs):
    """Integer points of Brush.find_smear_directions (brush.ml:239-283) for all the segments of a path in order."""
    rec = _seg_records(segs)
    cap = 4096
    while True:
        out = np.zeros((cap, 2), dtype=np.int32)
        n = lib().coh_host_smear_points(rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), _i32p(out), C.c_int64(cap))
        if n <= cap:
            return out[:n]
        cap = int(n)


def host_brush_points(segs, radius):
    rec = _seg_records(segs)
    cap = 4096
    while True:
        out = np.zeros((cap, 2), dtype=np.int32)
        n = lib().coh_host_brush_points(rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), C.c_double(radius), _i32p(out), C.c_int64(cap))
        if n <= cap:
            return out[:n]
        cap = int(n)
