"""Host-side scene construction, mirroring the reference's scene types (render.ml:19-75)
and the geometry preparation that happens before the raster hot path
(polygon.ml:83-127,222-287 flattening; coord.ml:44-50; brush.ml:126-130,172;
render.ml:556-586 primitives; render.ml:1485-1487 colours).

The device input boundary is AFTER the affine transform: integer edges in sub-pixel bins.
Python floats are IEEE binary64 with one rounding per operation, like OCaml's, so the
host arithmetic here follows the reference's operation order literally.
"""
import ctypes as C
import json
import math
import os

import numpy as np

from .abi import (COH_EVENODD, COH_FILL_AXIAL, COH_FILL_EXT_E, COH_FILL_EXT_S, COH_FILL_PLAIN, COH_FILL_RADIAL,
                  COH_NONZERO, COH_OBJ_BRUSH, COH_OBJ_GROUP_BEGIN, COH_OBJ_GROUP_END, COH_OBJ_PATH,
                  COH_OBJ_PRIMITIVE, CohObject)

CURVE_ACCURACY = 0.2  # polygon.ml:19


# ---- coord.ml ---------------------------------------------------------------------
def tdiv(a, b):
    """OCaml integer division: truncation toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def pix_of_sub(n):  # coord.ml:44
    return tdiv(n + 31, 32)


def sub_of_float(f):  # coord.ml:47
    return int(math.ceil(f * 32.0 - 16.0))


def pix_of_float(f):  # coord.ml:50
    return pix_of_sub(sub_of_float(f))


# ---- colours (colour.ml:247-252, 291-304) ------------------------------------------
def div255(i):
    return (i + (i >> 8) + 1) >> 8


def rgba8(r, g, b, a=255):
    return (r & 255) | ((g & 255) << 8) | ((b & 255) << 16) | ((a & 255) << 24)


def colour_of_rgba_float(r, g, b, a=1.0):
    return rgba8(int(r * 255.0), int(g * 255.0), int(b * 255.0), int(a * 255.0))


def dissolve(c, delta):
    if delta == 0:
        return 0
    if delta == 255:
        return c
    ch = [(c >> s) & 255 for s in (0, 8, 16, 24)]
    return rgba8(*[div255(v * delta) for v in ch])


LIGHTGREY = rgba8(211, 211, 211)  # colour.ml:482
WHITE = rgba8(255, 255, 255)
BLACK = rgba8(0, 0, 0)


# ---- path flattening (polygon.ml:83-127) ---------------------------------------------
def _distance_point_from_line(c, a, b):
    l = math.sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]))
    try:
        s = ((a[1] - c[1]) * (b[0] - a[0]) - (a[0] - c[0]) * (b[1] - a[1])) / (l * l)
    except ZeroDivisionError:  # OCaml yields nan/inf; classify_float then says "not normal" -> flat
        return float("nan")
    return abs(s) * l


def _fp_normal(d):
    return not (math.isnan(d) or math.isinf(d) or d == 0.0 or abs(d) < 2.2250738585072014e-308)


def bezier_epsilon(eps, p1, p2, p3, p4):
    d1 = _distance_point_from_line(p2, p1, p4)
    d2 = _distance_point_from_line(p3, p1, p4)
    if _fp_normal(d1) and _fp_normal(d2):
        return d1 < eps and d2 < eps
    return True


def bezier_subdivide(eps, p1, p2, p3, p4, out):
    if bezier_epsilon(eps, p1, p2, p3, p4):
        out.append((p1, p4))
        return
    half = lambda a, b: ((a[0] + b[0]) / 2.0, (a[1] + b[1]) / 2.0)
    l2, h = half(p1, p2), half(p2, p3)
    l3, r3 = half(l2, h), half(p3, p4)
    r2 = half(h, r3)
    l4 = half(l3, r2)
    bezier_subdivide(eps, p1, l2, l3, l4, out)
    bezier_subdivide(eps, l4, r2, r3, p4, out)


def flatten_segments(segs, eps=CURVE_ACCURACY):
    """segs: list of ('L', p1, p2) | ('C', p1, p2, p3, p4) -> list of (p, q) float pairs."""
    out = []
    for s in segs:
        if s[0] == "L":
            out.append((s[1], s[2]))
        else:
            bezier_subdivide(eps, s[1], s[2], s[3], s[4], out)
    return out


def edges_of_float_segments(pairs):
    """polygon.ml:262-287: every endpoint through Coord.sub_of_float."""
    return [[sub_of_float(p[0]), sub_of_float(p[1]), sub_of_float(q[0]), sub_of_float(q[1])] for p, q in pairs]


def polygon_segments(points):
    """Closed polygon through `points` (polygon.ml:66-76 path_of_pointlist)."""
    pts = list(points) + [points[0]]
    return [("L", pts[i], pts[i + 1]) for i in range(len(pts) - 1)]


def bounds_polygon(pairs):
    """polygon.ml:405-440 for straight segments: pix_of_float of every endpoint."""
    xs = [pix_of_float(p[0]) for pq in pairs for p in pq]
    ys = [pix_of_float(p[1]) for pq in pairs for p in pq]
    return min(xs), max(xs), min(ys), max(ys)


# ---- brush points (polygon.ml:143-218, brush.ml:25-28,126-130,172) --------------------
def points_on_path(sep, subpaths):
    points = []
    for segs in subpaths:
        work = []
        for s in segs:
            if s[0] == "L":
                work.insert(0, (s[1], s[2]))
            else:
                e = []
                bezier_subdivide(CURVE_ACCURACY, s[1], s[2], s[3], s[4], e)
                work[0:0] = e
        i = 0
        while i < len(work):
            want, found = sep, False
            while i < len(work):
                p1, p2 = work[i]
                l = math.sqrt((p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]))
                if want <= l:
                    prop = want / l
                    p = (p1[0] * (1.0 - prop) + p2[0] * prop, p1[1] * (1.0 - prop) + p2[1] * prop)
                    points.append(p)
                    if p == p2:
                        i += 1
                    else:
                        work[i] = (p, p2)
                    found = True
                    break
                want -= l
                i += 1
            if not found:
                break
    return points


def brush_points(radius, subpaths):
    w = int(math.ceil(radius)) * 2 + 1
    pts = points_on_path(float(w) / 20.0, subpaths)
    return [(int(x + 0.5), int(y + 0.5)) for x, y in pts]


# ---- fills (fill.ml:62-140) ------------------------------------------------------------
class Fill:
    def __init__(self, kind, c0, c1=0, flags=0, params=()):
        self.kind, self.c0, self.c1, self.flags = kind, c0, c1, flags
        self.params = list(params) + [0.0] * (6 - len(params))

    @staticmethod
    def plain(colour):
        return Fill(COH_FILL_PLAIN, colour)

    @staticmethod
    def gradient(p0, p1, ext_s, ext_e, cs, ce):
        return Fill(COH_FILL_AXIAL, cs, ce, (COH_FILL_EXT_S if ext_s else 0) | (COH_FILL_EXT_E if ext_e else 0), [p0[0], p0[1], p1[0], p1[1]])

    @staticmethod
    def radial(c, p, p2, ext_s, ext_e, cs, ce):
        return Fill(COH_FILL_RADIAL, cs, ce, (COH_FILL_EXT_S if ext_s else 0) | (COH_FILL_EXT_E if ext_e else 0), [c[0], c[1], p[0], p[1], p2[0], p2[1]])

    def apply(self, o):
        o.fill_kind, o.colour0, o.colour1, o.fill_flags = self.kind, self.c0, self.c1, self.flags
        for i in range(6):
            o.fparam[i] = self.params[i]


# ---- scene builder ---------------------------------------------------------------------
class SceneBuilder:
    """Flattens render.ml `renderobject`s (head = front-most) into the C-ABI arrays."""

    def __init__(self):
        self.objs, self.edges, self.points = [], [], []
        self._n_edges = self._n_points = 0
        self.n_background = 0
        self._in_background = False

    def _obj(self, kind, pretrans=None, oid=-1, dx=0, dy=0, convolve=None):
        """convolve = ("unit" | "gaussian", r): Convolved (Convolve.mkunit r | mkgaussian r, this geometry) — a path, or a
        group (group_begin(convolve=...): Convolved (kernel, Group members))."""
        o = CohObject()
        o.kind, o.pretrans, o.id, o.dx, o.dy = kind, (-1 if pretrans is None else pretrans), oid, dx, dy
        if convolve is not None:
            assert kind in (COH_OBJ_PATH, COH_OBJ_GROUP_BEGIN) and convolve[1] > 0
            o.convolve = {"unit": 1, "gaussian": 2}[convolve[0]] | (int(convolve[1]) << 8)
        self.objs.append(o)
        if self._in_background:
            self.n_background += 1
        return o

    def begin_background(self):
        """Everything added from here on is render_frame's (pages @ background) list."""
        self._in_background = True

    def stroked_path_edges(self, edges, fill, **kw):
        """Basic (fill, StrokedPath (p, spec)) with edges = Shapes.strokepath spec p (the stroker itself is host
        geometry outside this path): shape by NonZero (render.ml:510), sprite by EvenOdd (render.ml:1018)."""
        o = self.path_edges(edges, fill, COH_NONZERO, **kw)
        o.sprite_winding = 1 + COH_EVENODD
        return o

    def path_edges(self, edges, fill, winding=COH_NONZERO, bounds=None, **kw):
        """Basic (fill, Path p) with p already flattened to integer sub-bin edges."""
        o = self._obj(COH_OBJ_PATH, **kw)
        e = np.asarray(edges, dtype=np.int32).reshape(-1, 4)
        o.winding, o.first, o.count = winding, self._n_edges, len(e)
        self.edges.append(e)
        self._n_edges += len(e)
        fill.apply(o)
        if bounds is None and len(e):
            # bounds_polygon (polygon.ml:405-440): pix_of_float of every end point = pix_of_sub of the edge ends
            t = e.astype(np.int64) + 31
            pix = np.where(t >= 0, t // 32, -((-t) // 32))  # truncating division (coord.ml:44)
            bounds = (int(pix[:, [0, 2]].min()), int(pix[:, [0, 2]].max()), int(pix[:, [1, 3]].min()), int(pix[:, [1, 3]].max()))
        if bounds:
            for i in range(4):
                o.bounds[i] = bounds[i]
        return o

    def path(self, subpaths, fill, winding=COH_NONZERO, **kw):
        """subpaths: list of segment lists (device-space floats); flattened by the library's host-side
        geometry (Polygon.edgelist_of_path)."""
        from . import abi

        parts = [abi.host_edgelist_of_subpath(segs) for segs in subpaths]
        return self.path_edges(np.concatenate(parts) if parts else np.zeros((0, 4), np.int32), fill, winding, **kw)

    def polygon(self, points, fill, winding=COH_NONZERO, **kw):
        return self.path([polygon_segments(points)], fill, winding, **kw)

    def rectangle(self, colour, xmin, ymin, xmax, ymax, **kw):
        """Primitive (colour, Rectangle (xmin, ymin, xmax, ymax)) — render.ml:573-586: inclusive of toint max."""
        assert xmax >= xmin and ymax >= ymin
        o = self._obj(COH_OBJ_PRIMITIVE, **kw)
        o.colour0 = colour
        box = (int(xmin), int(ymin), int(xmax), int(ymax))
        for i in range(4):
            o.prim[i] = box[i]
        for i, v in enumerate((box[0], box[2], box[1], box[3])):
            o.bounds[i] = v
        return o

    def hline(self, colour, y, xmin, xmax, **kw):
        o = self._obj(COH_OBJ_PRIMITIVE, **kw)
        o.colour0 = colour
        yi, x0, x1 = int(y), int(xmin), int(xmax)
        assert x1 >= x0
        o.prim_null = 1 if x1 == x0 else 0  # render.ml:560
        for i, v in enumerate((x0, yi, x1, yi)):
            o.prim[i] = v
        return o

    def vline(self, colour, x, ymin, ymax, **kw):
        o = self._obj(COH_OBJ_PRIMITIVE, **kw)
        o.colour0 = colour
        xi, y0, y1 = int(x), int(ymin), int(ymax)
        assert y1 >= y0
        o.prim_null = 1 if y1 == y0 else 0  # render.ml:567
        for i, v in enumerate((xi, y0, xi, y1)):
            o.prim[i] = v
        return o

    def brush(self, opacity, radius, subpaths, fill, **kw):
        """Basic (fill, Brushstroke ((opacity, Gaussian radius), path))."""
        from . import abi

        o = self._obj(COH_OBJ_BRUSH, **kw)
        parts = [abi.host_brush_points(segs, radius) for segs in subpaths]
        pts = np.concatenate(parts) if parts else np.zeros((0, 2), np.int32)
        o.first, o.count = self._n_points, len(pts)
        self.points.append(pts)
        self._n_points += len(pts)
        o.brush_opacity, o.brush_radius = float(opacity), float(radius)
        fill.apply(o)
        return o

    def dummy_brush(self, radius, subpaths, **kw):
        """Basic (_, Brushstroke (Brush.mkdummy ((opacity, Gaussian radius), path))) (brush.ml:70-73): the stroke's whole
        shape in white."""
        import math

        from . import abi

        o = self.brush(1.0, radius, subpaths, Fill.plain(WHITE), **kw)
        o.winding = abi.COH_BRUSH_DUMMY
        o.brush_radius = float(math.ceil(radius))   # ((2 ceil r + 1) - 1) / 2
        return o

    def cpg(self, op, subpaths_a, subpaths_b, fill, winding_a=COH_NONZERO, winding_b=COH_NONZERO, **kw):
        """Basic (fill, CPG (op, Path a, Path b)); op in "union" | "intersection" | "subtraction" | "xor"."""
        from . import abi

        o = self._obj(abi.COH_OBJ_CPG, **kw)
        ea = np.concatenate([abi.host_edgelist_of_subpath(sg) for sg in subpaths_a]).reshape(-1, 4)
        eb = np.concatenate([abi.host_edgelist_of_subpath(sg) for sg in subpaths_b]).reshape(-1, 4)
        o.first, o.count, o.winding = self._n_edges, len(ea), winding_a
        o.first2, o.count2, o.winding2 = self._n_edges + len(ea), len(eb), winding_b
        o.cpg_op = {"union": 0, "intersection": 1, "subtraction": 2, "xor": 3}[op]
        self.edges.extend([ea, eb])
        self._n_edges += len(ea) + len(eb)
        fill.apply(o)
        return o

    def filter(self, kind, subpaths, fill=None, winding=COH_NONZERO, kernel=None, **kw):
        """Filter {geometry = Basic (fill, Path subpaths); ...} (filters.ml): kind in "hole" | "monochrome" | "blur" |
        "scene".  kernel = ("gaussian" | "unit", r) for blur; a "scene" filter (affine, rgb, wireframe ...: the
        caller rewrites the objects below) gets its modified scene from reading_scene_begin(filter_obj)."""
        from . import abi

        o = self._obj(abi.COH_OBJ_FILTER, **kw)
        e = np.concatenate([abi.host_edgelist_of_subpath(sg) for sg in subpaths]).reshape(-1, 4)
        o.first, o.count, o.winding = self._n_edges, len(e), winding
        self.edges.append(e)
        self._n_edges += len(e)
        (fill or Fill.plain(WHITE)).apply(o)
        o.filter_kind = {"hole": 1, "monochrome": 2, "blur": 3, "scene": 4, "minus": 5}[kind]
        if kind == "blur":
            o.filter_kernel = {"unit": 1, "gaussian": 2}[kernel[0]] | (int(kernel[1]) << 8)
        return o

    def filter_with_geometry(self, kind, kernel=None, **kw):
        """A filter whose geometry is the OBJECT ADDED NEXT (a brush stroke, a Convolved path, a CPG, a group ...): that
        object is consumed by the filter, its shape is the filter's and the alpha of its sprite the matte."""
        from . import abi

        o = self._obj(abi.COH_OBJ_FILTER, **kw)
        o.cpg_op = abi.COH_GEOM_NEXT
        Fill.plain(WHITE).apply(o)
        o.filter_kind = {"hole": 1, "monochrome": 2, "blur": 3, "scene": 4, "minus": 5}[kind]
        if kind == "blur":
            o.filter_kernel = {"unit": 1, "gaussian": 2}[kernel[0]] | (int(kernel[1]) << 8)
        return o

    def smear_filter(self, opacity, radius, subpaths, **kw):
        """Filters.smear ((opacity, Gaussian radius), path) (filters.ml:201-217): geometry = the stroke's dummy brush
        (its stamp points), plus the integer smear points of Brush.find_smear_directions; both in the points array."""
        from . import abi

        o = self._obj(abi.COH_OBJ_FILTER, **kw)
        parts = [abi.host_brush_points(segs, radius) for segs in subpaths]
        pts = np.concatenate(parts) if parts else np.zeros((0, 2), np.int32)
        o.first, o.count = self._n_points, len(pts)
        self.points.append(pts)
        self._n_points += len(pts)
        sm = abi.host_smear_points([sg for segs in subpaths for sg in segs])
        o.first2, o.count2 = self._n_points, len(sm)
        self.points.append(sm)
        self._n_points += len(sm)
        o.brush_opacity, o.brush_radius = float(opacity), float(radius)
        Fill.plain(WHITE).apply(o)
        o.filter_kind = abi.COH_FILTER_SMEAR
        return o

    def reading_scene_begin(self, filter_obj):
        """Open the reading-scene group of a "scene" filter (after every ordinary scene object); close with group_end()."""
        from . import abi

        o = self.group_begin()
        o.filter_kind = abi.COH_FILTER_READING_SCENE
        filter_obj.first2 = len(self.objs) - 1
        return o

    def group_begin(self, **kw):
        return self._obj(COH_OBJ_GROUP_BEGIN, **kw)

    def group_end(self):
        return self._obj(COH_OBJ_GROUP_END)

    def arrays(self):
        arr = (CohObject * max(len(self.objs), 1))(*self.objs)
        e = np.concatenate(self.edges).reshape(-1, 4) if self.edges else np.zeros((0, 4), np.int32)
        p = np.concatenate(self.points).reshape(-1, 2) if self.points else np.zeros((0, 2), np.int32)
        return arr, len(self.objs), self.n_background, e, p


# ---- benchmark scenes ---------------------------------------------------
_SCENES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes")


def lion_paths():
    with open(os.path.join(_SCENES, "lion.json")) as f:
        return json.load(f)["paths"]


def add_lion(b, width, height, scale, pretrans=None, oid=-1, transform=None):
    """Append the lion (examples.ml:174-180: Group (rev objs), Over), centred on a width x height canvas, y flipped
    about its bounding box, uniformly scaled; `transform` maps device-space points (an affine filter's rewrite)."""
    paths = lion_paths()
    xs = [p[0] for q in paths for p in q["subpaths"][0]]
    ys = [p[1] for q in paths for p in q["subpaths"][0]]
    cx, cy = (min(xs) + max(xs)) / 2.0, (min(ys) + max(ys)) / 2.0

    def tr(p):
        q = ((p[0] - cx) * scale + width / 2.0, (cy - p[1]) * scale + height / 2.0)
        return transform(q) if transform else q

    g = b.group_begin(pretrans=pretrans, oid=oid)
    for q in reversed(paths):  # head = front-most = painted last
        r, g8, bl = q["rgb"]
        fill = Fill.plain(colour_of_rgba_float(r, g8, bl, 1.0))
        subs = []
        for sp in q["subpaths"]:
            pts = [tr(p) for p in sp]
            subs.append([("L", pts[i], pts[i + 1]) for i in range(len(pts) - 1)])
        b.path(subs, fill, COH_NONZERO)
    b.group_end()
    return g


def add_text_page(b, scale, origin=(0.0, 0.0), name="mintext1", **group_kw):
    """Append a text page of the reference (examples.ml:158-165 load_text: Group (rev objs)) from its committed geometry
    (scenes/<name>.json, written by tools/make_text_fixture.py through pdf_import): glyph outlines as filled paths of
    lines and curves.  x' = ox + scale x, y' = oy + scale (page height - y)."""
    with open(os.path.join(_SCENES, name + ".json")) as f:
        page = json.load(f)
    ox, oy = origin
    ph = page["mediabox"][3]

    def tr(x, y):
        return (ox + scale * x, oy + scale * (ph - y))

    g = b.group_begin(**group_kw)
    for q in reversed(page["paths"]):
        subs = []
        for sp in q["subpaths"]:
            subs.append([(("L" if len(s) == 4 else "C"),) + tuple(tr(s[i], s[i + 1]) for i in range(0, len(s), 2)) for s in sp])
        b.path(subs, Fill.plain(q["colour"]), COH_NONZERO if q["winding"] == "NonZero" else COH_EVENODD)
    b.group_end()
    return g


def lion_scene(width, height, scale, background=True, pretrans=None):
    """C1/C2: the lion over Primitive (lightgrey, Rectangle (0, 0, W, H)) (engine.ml:73-74)."""
    b = SceneBuilder()
    add_lion(b, width, height, scale, pretrans=pretrans)
    if background:
        b.begin_background()
        b.rectangle(LIGHTGREY, 0.0, 0.0, float(width), float(height))
    return b


def circle_subpath(cx, cy, r, n=64):
    return [polygon_segments([(cx + r * math.cos(2.0 * math.pi * i / n), cy + r * math.sin(2.0 * math.pi * i / n)) for i in range(n)])]


def filter_scene(width, height, scale):
    """C5: the lion under a blur lens (circle, mkgaussian 5), a monochrome lens (circle), an affine lens
    (rectangle; reading scene = the lion squashed and sheared about the lens centre, filters.ml:271-285) and a page
    shadow Convolved (mkgaussian 4, rectangle) (engine.ml:85-88), laid out like examples.ml:69-88."""
    W, H = float(width), float(height)
    u = min(W, H)
    b = SceneBuilder()
    b.filter("blur", circle_subpath(0.36 * W, 0.40 * H, 0.185 * u), kernel=("gaussian", 5))
    b.filter("monochrome", circle_subpath(0.62 * W, 0.36 * H, 0.16 * u))
    ax0, ay0, ax1, ay1 = 0.42 * W, 0.58 * H, 0.70 * W, 0.88 * H
    acx, acy = (ax0 + ax1) / 2.0, (ay0 + ay1) / 2.0
    aff = b.filter("scene", [polygon_segments([(ax0, ay0), (ax1, ay0), (ax1, ay1), (ax0, ay1)])])
    add_lion(b, width, height, scale)
    shadow = [polygon_segments([(0.08 * W, 0.08 * H), (0.92 * W, 0.08 * H), (0.92 * W, 0.92 * H), (0.08 * W, 0.92 * H)])]
    b.path(shadow, Fill.plain(dissolve(rgba8(0, 0, 0), 120)), COH_NONZERO, convolve=("gaussian", 4))

    def affine(p):  # Scale ((cx, cy), 1, -0.5) then ShearX ((cx, cy), -0.3)
        x, y = p[0], acy + (p[1] - acy) * -0.5
        return (x + (y - acy) * -0.3, y)

    b.reading_scene_begin(aff)
    add_lion(b, width, height, scale, transform=affine)
    b.group_end()
    b.begin_background()
    b.rectangle(LIGHTGREY, 0.0, 0.0, W, H)
    return b


def drag_scene(width, height, scale, n_static=400, seed=0xD1CE):
    """C4: n_static C3-style objects with the lion group (id 1) in front as the dragged object.
    Returns (builder, abi index of the lion group)."""
    b = SceneBuilder()
    add_lion(b, width, height, scale, oid=1)
    static = random_scene(width, height, n_static, seed=seed, brush_fraction=0.0, background=False)
    for k, o in enumerate(static.objs):
        o.first += b._n_edges
        o.id = 1000 + k
        b.objs.append(o)
    b.edges.extend(static.edges)
    b._n_edges += static._n_edges
    b.begin_background()
    b.rectangle(LIGHTGREY, 0.0, 0.0, float(width), float(height))
    return b, 0


class PCG32:
    """Minimal PCG-XSH-RR 64/32 (O'Neill 2014) for reproducible synthetic scenes."""

    def __init__(self, seed, seq=54):
        self.state, self.inc = 0, ((seq << 1) | 1) & 0xFFFFFFFFFFFFFFFF
        self.next()
        self.state = (self.state + seed) & 0xFFFFFFFFFFFFFFFF
        self.next()

    def next(self):
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & 0xFFFFFFFFFFFFFFFF
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform(self):
        return self.next() / 4294967296.0

    def randint(self, lo, hi):  # inclusive
        return lo + self.next() % (hi - lo + 1)


def random_scene(width, height, n_objects, seed=0xC0FFEE, brush_fraction=0.2, background=True):
    """C3: random layered polygons / brush strokes."""
    rng = PCG32(seed)
    b = SceneBuilder()
    for _ in range(n_objects):
        is_brush = rng.uniform() < brush_fraction
        r8, g8, b8 = rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255)
        alpha = 255 if rng.uniform() < 0.7 else rng.randint(64, 254)
        fill = Fill.plain(dissolve(rgba8(r8, g8, b8, 255), alpha))
        cxp, cyp = rng.uniform() * width, rng.uniform() * height
        if not is_brush:
            k = rng.randint(3, 12)
            radius = math.exp(math.log(8.0) + rng.uniform() * (math.log(400.0) - math.log(8.0)))
            angles = sorted(rng.uniform() * 2.0 * math.pi for _ in range(k))
            pts = []
            for a in angles:
                rr = radius * (0.6 + 0.4 * rng.uniform())
                pts.append((cxp + rr * math.cos(a), cyp + rr * math.sin(a)))
            b.polygon(pts, fill, COH_NONZERO)
        else:
            rad = float(rng.randint(3, 20))
            opacity = 0.5 + 0.5 * rng.uniform()
            cps = [(cxp + (rng.uniform() - 0.5) * 600.0, cyp + (rng.uniform() - 0.5) * 600.0) for _ in range(4)]
            b.brush(opacity, rad, [[("C", cps[0], cps[1], cps[2], cps[3])]], fill)
    if background:
        b.begin_background()
        b.rectangle(LIGHTGREY, 0.0, 0.0, float(width), float(height))
    return b
