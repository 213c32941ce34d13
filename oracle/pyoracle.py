"""ctypes binding of the ORACLE (CPU restatement of the reference renderer).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED: the reference has no tests or golden vectors and cannot be built here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_rgba8_of_colour.restype = C.c_uint32
        _LIB.orc_renderer_new.restype = C.c_void_p
    return _LIB


class OracleError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, t=C.c_int32):
    return a.ctypes.data_as(C.POINTER(t))


def _take(p, n):
    out = np.ctypeslib.as_array(p, shape=(max(int(n), 1),))[: int(n)].copy()
    lib().orc_free(p)
    return out


def smear_points(seg_records):
    """Brush.points_of_brushstroke_smear + integer points of find_smear_directions; seg_records: float64 [n][9]."""
    rec = np.ascontiguousarray(seg_records, dtype=np.float64).reshape(-1, 9)
    lib().orc_smear_points.restype = C.c_int64
    cap = 1 << 16
    out = np.zeros((cap, 2), dtype=np.int32)
    n = lib().orc_smear_points(rec.ctypes.data_as(C.POINTER(C.c_double)), len(rec), out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(cap))
    return out[:n]


def colour_of_rgba8(w):
    return lib().orc_colour_of_rgba8(C.c_uint32(w))


def rgba8_of_colour(c):
    return lib().orc_rgba8_of_colour(C.c_int32(c))


def colour_of_rgba(r, g, b, a):
    return lib().orc_colour_of_rgba(r, g, b, a)


def colour_op(op, a, b=0, c=0):
    ops = {"over": 0, "alpha_over": 1, "pd_plus": 2, "dissolve": 3, "dissolve_between": 4, "monochrome": 5}
    out = C.c_uint32()
    _chk(lib().orc_colour_op(ops[op], C.c_uint32(a), C.c_uint32(b), c, C.byref(out)))
    return out.value


def div255(i):
    return lib().orc_div255(i)


def sub_of_float(f):
    return lib().orc_sub_of_float(C.c_double(f))


def pix_of_sub(n):
    return lib().orc_pix_of_sub(n)


def aa_tables():
    t = np.zeros(32 * 32, dtype=np.int32)
    v = C.c_int32()
    lib().orc_aa_tables(_ptr(t), C.byref(v))
    return t.reshape(32, 32), v.value


def shapeminshape(edges, winding):
    e = _i32(edges).reshape(-1, 4)
    ps, pm = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    ns, nm = C.c_int64(), C.c_int64()
    _chk(lib().orc_shapeminshape(_ptr(e), len(e), winding, C.byref(ps), C.byref(ns), C.byref(pm), C.byref(nm)))
    return _take(ps, ns.value), _take(pm, nm.value)


def scaled_shape(edges, winding):
    e = _i32(edges).reshape(-1, 4)
    ps, ns = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_scaled_shape(_ptr(e), len(e), winding, C.byref(ps), C.byref(ns)))
    return _take(ps, ns.value)


def shape_card(flat):
    n, i, flat = 0, 0, np.asarray(flat)
    while i < len(flat):
        k = int(flat[i + 1])
        n += int(flat[i + 3 : i + 2 + 2 * k : 2].sum())
        i += 2 + 2 * k
    return n


def polygon_opacity(edges, winding, shape_flat):
    e = _i32(edges).reshape(-1, 4)
    s = _i32(shape_flat)
    cap = shape_card(s)
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    n = C.c_int64()
    _chk(lib().orc_polygon_opacity(_ptr(e), len(e), winding, _ptr(s), C.c_int64(len(s)), _ptr(out, C.c_uint8), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


def polygon_sprite(fill_obj, edges, winding, shape_flat):
    e = _i32(edges).reshape(-1, 4)
    s = _i32(shape_flat)
    cap = shape_card(s)
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_int64()
    _chk(lib().orc_polygon_sprite(C.byref(fill_obj), _ptr(e), len(e), winding, _ptr(s), C.c_int64(len(s)), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


def shape_op(op, a, b):
    ops = {"union": 0, "difference": 1, "intersection": 2}
    a, b = _i32(a), _i32(b)
    p, n = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_shape_op(ops[op], _ptr(a), C.c_int64(len(a)), _ptr(b), C.c_int64(len(b)), C.byref(p), C.byref(n)))
    return _take(p, n.value)


def shape_unary(op, a, m, n_):
    ops = {"bloat": 0, "erode": 1, "translate": 2}
    a = _i32(a)
    p, n = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_shape_unary(ops[op], _ptr(a), C.c_int64(len(a)), m, n_, C.byref(p), C.byref(n)))
    return _take(p, n.value)


def render_frame(objs, n_scene, n_background, edges, points, update, bbox_reject=True, usecache=False, want_u=False):
    """objs: ctypes array of coh_object (scene objects then background objects)."""
    ux, uy, uw, uh = update
    e = _i32(edges).reshape(-1, 4)
    p = _i32(points).reshape(-1, 2)
    out = np.zeros((uh, uw), dtype=np.uint32)
    pu, nu = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(
        lib().orc_render_frame(
            objs, n_scene, n_background, _ptr(e), _ptr(p), ux, uy, uw, uh, 0 if bbox_reject else 1, 1 if usecache else 0,
            _ptr(out, C.c_uint32), C.byref(pu) if want_u else None, C.byref(nu) if want_u else None,
        )
    )
    if want_u:
        return out, _take(pu, nu.value)
    return out


def convolve_sprite(kernel, r, shape_flat, rgba):
    s = _i32(shape_flat)
    src = np.ascontiguousarray(rgba, dtype=np.uint32)
    cap = shape_card(shape_unary("bloat", s, r, r))
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    ps, ns, n = C.POINTER(C.c_int32)(), C.c_int64(), C.c_int64()
    _chk(lib().orc_convolve_sprite({"unit": 1, "gaussian": 2}[kernel], r, _ptr(s), C.c_int64(len(s)), _ptr(src, C.c_uint32), C.byref(ps), C.byref(ns), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return _take(ps, ns.value), out[: n.value]


def brush_shape(brush_obj, points):
    p = _i32(points).reshape(-1, 2)
    ps, ns = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_brush_shape(C.byref(brush_obj), _ptr(p), len(p), C.byref(ps), C.byref(ns)))
    return _take(ps, ns.value)


def brush_sprite(brush_obj, points, shape_flat):
    p = _i32(points).reshape(-1, 2)
    s = _i32(shape_flat)
    cap = shape_card(s)
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_int64()
    _chk(lib().orc_brush_sprite(C.byref(brush_obj), _ptr(p), len(p), _ptr(s), C.c_int64(len(s)), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


def brush_smear(shape_flat, rgba, brush_obj, points, smear_pts):
    """Brush.smear: returns (flat shape of the result, rgba per pixel)."""
    s = _i32(shape_flat)
    src = np.ascontiguousarray(rgba, dtype=np.uint32)
    p = _i32(points).reshape(-1, 2)
    q = _i32(smear_pts).reshape(-1, 2)
    cap = shape_card(shape_op("union", s, brush_shape(brush_obj, p))) if len(s) else shape_card(brush_shape(brush_obj, p))
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    ps, ns, n = C.POINTER(C.c_int32)(), C.c_int64(), C.c_int64()
    _chk(lib().orc_brush_smear(_ptr(s), C.c_int64(len(s)), _ptr(src, C.c_uint32), C.byref(brush_obj), _ptr(p), len(p), _ptr(q), len(q), C.byref(ps), C.byref(ns),
                               _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return _take(ps, ns.value), out[: n.value]


def flatten_bezier(p8, eps=0.2):
    a = np.ascontiguousarray(p8, dtype=np.float64).reshape(8)
    p, n = C.POINTER(C.c_double)(), C.c_int64()
    _chk(lib().orc_flatten_bezier(_ptr(a, C.c_double), C.c_double(eps), C.byref(p), C.byref(n)))
    out = np.ctypeslib.as_array(p, shape=(max(n.value, 1) * 4,))[: n.value * 4].copy().reshape(-1, 4)
    lib().orc_free(p)
    return out


def points_on_path(segs, sep):
    a = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 9)
    p, n = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_points_on_path(_ptr(a, C.c_double), len(a), C.c_double(sep), C.byref(p), C.byref(n)))
    return _take(p, n.value * 2).reshape(-1, 2)


def strokepath(spec, segs, subpath_n):
    """Shapes.strokepath_polygon + Shapes.strokepath.  spec = (startcap, join, endcap, mitrelimit, linewidth) with the
    oracle's enums (caps Butt 0 / Round 1 / Projecting 2; joins Round 0 / Mitred 1 / Bevel 2).
    Returns (outline segments (n, 9), segments per outline subpath, winding, sorted edges (m, 4))."""
    a = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 9)
    cnt = np.ascontiguousarray(subpath_n, dtype=np.int32)
    sp = np.ascontiguousarray(spec, dtype=np.float64)
    so, ns, co, nc, w = C.POINTER(C.c_double)(), C.c_int64(), C.POINTER(C.c_int32)(), C.c_int64(), C.c_int()
    eo, ne = C.POINTER(C.c_int32)(), C.c_int64()
    _chk(lib().orc_strokepath(_ptr(sp, C.c_double), _ptr(a, C.c_double), _ptr(cnt, C.c_int32), len(cnt), C.byref(so), C.byref(ns),
                              C.byref(co), C.byref(nc), C.byref(w), C.byref(eo), C.byref(ne)))
    out = np.ctypeslib.as_array(so, shape=(max(ns.value * 9, 1),))[: ns.value * 9].copy().reshape(-1, 9)
    lib().orc_free(so)
    return out, _take(co, nc.value), w.value, _take(eo, ne.value * 4).reshape(-1, 4)


def bounds_stroke(spec, segs, subpath_n):
    """Shapes.bounds_stroke: (xmin, xmax, ymin, ymax) in pixels."""
    a = np.ascontiguousarray(segs, dtype=np.float64).reshape(-1, 9)
    cnt = np.ascontiguousarray(subpath_n, dtype=np.int32)
    sp = np.ascontiguousarray(spec, dtype=np.float64)
    out = np.zeros(4, dtype=np.int32)
    _chk(lib().orc_bounds_stroke(_ptr(sp, C.c_double), _ptr(a, C.c_double), _ptr(cnt, C.c_int32), len(cnt), _ptr(out)))
    return tuple(int(v) for v in out)


def brush_stamp(radius, opacity):
    out = np.zeros(101 * 101, dtype=np.uint8)
    size = C.c_int()
    _chk(lib().orc_brush_stamp(C.c_double(radius), C.c_double(opacity), _ptr(out, C.c_uint8), len(out), C.byref(size)))
    s = size.value
    return out[: s * s].reshape(s, s)


class Renderer:
    """A persistent oracle renderer: Cache (shapes, partial sprites, aliases) survives between frames, as the
    reference's does between calls of Render.render_frame (cache.ml).  usecache = Cache.usecache."""

    def __init__(self, usecache=True):
        self._h = C.c_void_p(lib().orc_renderer_new(1 if usecache else 0))

    def close(self):
        if self._h:
            lib().orc_renderer_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def addtranslation(self, oid, target, dx, dy):
        _chk(lib().orc_renderer_addtranslation(self._h, C.c_int64(oid), C.c_int64(target), dx, dy))

    def frame(self, objs, n_scene, n_background, edges, points, update_flat, out_box):
        """Render.render_frame over the update shape (flat records); returns the dense RGBA8 image of out_box."""
        ox, oy, ow, oh = out_box
        e = _i32(edges).reshape(-1, 4)
        p = _i32(points).reshape(-1, 2)
        u = _i32(update_flat)
        out = np.zeros((oh, ow), dtype=np.uint32)
        _chk(lib().orc_renderer_frame(self._h, objs, n_scene, n_background, _ptr(e), _ptr(p), _ptr(u), C.c_int64(len(u)), ox, oy, ow, oh, _ptr(out, C.c_uint32)))
        return out


def sprite_portion(shape_flat, rgba, sub_flat):
    s, b = _i32(shape_flat), _i32(sub_flat)
    src = np.ascontiguousarray(rgba, dtype=np.uint32)
    cap = shape_card(b)
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_int64()
    _chk(lib().orc_sprite_portion(_ptr(s), C.c_int64(len(s)), _ptr(src, C.c_uint32), _ptr(b), C.c_int64(len(b)), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


def sprite_fillshape(fill_obj, shape_flat):
    s = _i32(shape_flat)
    cap = shape_card(s)
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_int64()
    _chk(lib().orc_sprite_fillshape(C.byref(fill_obj), _ptr(s), C.c_int64(len(s)), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


def sprite_map(op, rgba, arg=0):
    src = np.ascontiguousarray(rgba, dtype=np.uint32)
    out = np.zeros(max(len(src), 1), dtype=np.uint32)
    code = {"monochrome": 0, "dissolve": 1, "red_channel": 2, "green_channel": 3, "blue_channel": 4}[op]
    _chk(lib().orc_sprite_map(code, arg, _ptr(src, C.c_uint32), C.c_int64(len(src)), _ptr(out, C.c_uint32)))
    return out[: len(src)]


def sprite_map_coords_fill(fill_obj, shape_flat, rgba):
    s = _i32(shape_flat)
    src = np.ascontiguousarray(rgba, dtype=np.uint32)
    cap = shape_card(s)
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    n = C.c_int64()
    _chk(lib().orc_sprite_map_coords_fill(C.byref(fill_obj), _ptr(s), C.c_int64(len(s)), _ptr(src, C.c_uint32), _ptr(out, C.c_uint32), C.c_int64(cap), C.byref(n)))
    return out[: n.value]


# ---- camlpy.ml (oracle/camlpy.hpp): Python values as pycaml.py maps them (None, bool, int, bytes / str, list) ----
def _wire_tokens(m, kinds, values, offsets, blob):
    if m is None:
        kinds.append(1), values.append(0), offsets.append(0)
    elif isinstance(m, bool):
        kinds.append(4), values.append(int(m)), offsets.append(0)
    elif isinstance(m, int):
        kinds.append(2), values.append(m), offsets.append(0)
    elif isinstance(m, (bytes, bytearray, str)):
        b = m.encode("latin-1") if isinstance(m, str) else bytes(m)
        kinds.append(3), values.append(len(b)), offsets.append(len(blob))
        blob.extend(b)
    else:
        kinds.append(0), values.append(len(m)), offsets.append(0)
        for e in m:
            _wire_tokens(e, kinds, values, offsets, blob)


def wire_marshal(m):
    """Camlpy.marshall (camlpy.ml:77-82)."""
    kinds, values, offsets, blob = [], [], [], bytearray()
    _wire_tokens(m, kinds, values, offsets, blob)
    k, v, o = np.asarray(kinds, dtype=np.int32), np.asarray(values, dtype=np.int64), np.asarray(offsets, dtype=np.int64)
    sb = np.frombuffer(bytes(blob) + b"\0", dtype=np.uint8)
    size = C.c_int64()
    _chk(lib().orc_wire_marshal(_ptr(k), _ptr(v, C.c_int64), _ptr(o, C.c_int64), len(k), _ptr(sb, C.c_uint8), None, C.c_int64(0), C.byref(size)))
    out = np.zeros(max(size.value, 1), dtype=np.uint8)
    _chk(lib().orc_wire_marshal(_ptr(k), _ptr(v, C.c_int64), _ptr(o, C.c_int64), len(k), _ptr(sb, C.c_uint8), _ptr(out, C.c_uint8), C.c_int64(size.value), C.byref(size)))
    return out[: size.value].tobytes()


def wire_unmarshal(data):
    """Camlpy.unmarshall (camlpy.ml:106-124): None, (taken, value) or OracleError("Invalid_data")."""
    buf = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8)
    cap = len(data) + 1
    k, v, o = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    nt, taken, status = C.c_int(), C.c_int64(), C.c_int()
    _chk(lib().orc_wire_unmarshal(_ptr(buf, C.c_uint8), C.c_int64(len(data)), _ptr(k), _ptr(v, C.c_int64), _ptr(o, C.c_int64), cap,
                                  C.byref(nt), C.byref(taken), C.byref(status)))
    if status.value < 0:
        raise OracleError("Invalid_data")
    if status.value == 0:
        return None
    pos = [0]

    def build():
        i = pos[0]
        pos[0] += 1
        if k[i] == 1:
            return None
        if k[i] == 4:
            return bool(v[i])
        if k[i] == 2:
            return int(v[i])
        if k[i] == 3:
            return bytes(data[int(o[i]) : int(o[i]) + int(v[i])])
        return [build() for _ in range(int(v[i]))]

    return taken.value, build()
