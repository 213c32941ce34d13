"""Trust-building checks for the oracle: canonical form, set algebra against dense
bitmaps, bloat against brute-force dilation, and the span algorithm's frame against a per-pixel fold."""
import random

import numpy as np
import pytest

from coherence_renderer_b200 import scene as S
from tests import util


def _canonical(flat):
    rows = util.rows_of_flat(flat)
    ys = [y for y, _ in rows]
    assert ys == sorted(set(ys))
    for _, spans in rows:
        assert spans, "empty spanline"
        for i, (x, l) in enumerate(spans):
            assert l > 0
            if i:
                px, pl = spans[i - 1]
                assert x > px + pl, "spans overlap or abut"  # sprite.ml:201-239


def test_set_algebra_equals_bitmap_algebra(oracle):
    rng = random.Random(21)
    X0, Y0, W, H = -40, -40, 420, 300
    for _ in range(40):
        a, b = util.random_shape_flat(rng), util.random_shape_flat(rng, x0=10, y0=-20, w=320, h=220)
        A, B = util.bitmap_of_flat(a, X0, Y0, W, H), util.bitmap_of_flat(b, X0, Y0, W, H)
        for name, ref in (("union", A | B), ("difference", A & ~B), ("intersection", A & B)):
            got = oracle.shape_op(name, a, b)
            _canonical(got)
            assert np.array_equal(util.bitmap_of_flat(got, X0, Y0, W, H), ref), name
            assert np.array_equal(got, util.flat_of_bitmap(ref, X0, Y0)), name  # canonical form is unique


def test_bloat_is_dilation_and_erode_contained(oracle):
    rng = random.Random(22)
    X0, Y0, W, H = -60, -60, 460, 340
    for _ in range(25):
        a = util.random_shape_flat(rng)
        m, n = rng.randint(0, 7), rng.randint(0, 7)
        A = util.bitmap_of_flat(a, X0, Y0, W, H)
        ref = np.zeros_like(A)
        for dy in range(-n, n + 1):
            for dx in range(-m, m + 1):
                ref |= np.roll(np.roll(A, dy, axis=0), dx, axis=1)
        got = oracle.shape_unary("bloat", a, m, n)
        _canonical(got)
        assert np.array_equal(util.bitmap_of_flat(got, X0, Y0, W, H), ref)
        er = util.bitmap_of_flat(oracle.shape_unary("erode", a, m, n), X0, Y0, W, H)
        assert not (er & ~A).any()
        back = util.bitmap_of_flat(oracle.shape_unary("bloat", oracle.shape_unary("erode", a, m, n), m, n), X0, Y0, W, H)
        assert not (back & ~A).any()  # opening is contained in the set


def test_scan_converter_canonical_and_nested(oracle):
    rng = random.Random(23)
    for _ in range(200):
        e = util.random_polygon_edges(rng)
        shp, mshp = oracle.shapeminshape(e, rng.randint(0, 1))
        _canonical(shp)
        _canonical(mshp)
        assert len(oracle.shape_op("difference", mshp, shp)) == 0  # minshape ⊆ shape (polygon.ml:526)
    assert all(len(x) == 0 for x in oracle.shapeminshape(np.zeros((0, 4), np.int32), 0))  # polygon.ml:584


def _over(a, b):
    aa = a >> 24
    if aa == 0:
        return b
    if aa == 255:
        return a
    out = 0
    for s in (0, 8, 16, 24):
        p, q = (b >> s) & 255, (a >> s) & 255
        t = aa * p + 128
        out |= (p + q - (((t >> 8) + t) >> 8)) << s
    return out


def test_frame_equals_per_pixel_fold(oracle):
    """End of SURVEY.md §8a: for a flat scene the span algorithm (caf / hidden-surface subtraction) must equal,
    pixel by pixel, the front-to-back fold of the objects' own sprites."""
    W, H = 96, 72
    rng = random.Random(24)
    b = S.SceneBuilder()
    layers = []
    for i in range(9):
        pts = [(rng.uniform(-10, W + 10), rng.uniform(-10, H + 10)) for _ in range(rng.randint(3, 6))]
        col = S.dissolve(S.rgba8(rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255)), rng.choice([255, 255, 160, 90]))
        pre = rng.choice([None, None, 200])
        o = b.polygon(pts, S.Fill.plain(col), rng.randint(0, 1), pretrans=pre)
        layers.append((o, col, pre))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    frame = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    fold = np.zeros((H, W), dtype=np.uint64)
    box = util.flat_of_rows([(y, [(0, W)]) for y in range(H)])
    for o, col, pre in layers:
        e = edges[o.first : o.first + o.count]
        shp, mshp = oracle.shapeminshape(e, o.winding)
        shp = oracle.shape_op("intersection", shp, box)
        if len(shp) == 0:
            continue
        op = oracle.polygon_opacity(e, o.winding, shp)
        inner = util.bitmap_of_flat(mshp, 0, 0, W, H)
        k = 0
        for y, spans in util.rows_of_flat(shp):
            for x, l in spans:
                for i in range(l):
                    c = col if inner[y, x + i] else S.dissolve(col, int(op[k]))
                    if pre is not None:
                        c = S.dissolve(c, pre)
                    fold[y, x + i] = _over(int(fold[y, x + i]), c)
                    k += 1
    for y in range(H):
        for x in range(W):
            fold[y, x] = _over(int(fold[y, x]), S.WHITE)
    assert np.array_equal(fold.astype(np.uint32), frame)


def test_caf_nocover_raises_on_overlap_via_cache_path(oracle):
    """spriteof unions the cached and the newly rendered parts with Colour.nocover (render.ml:1213,1231):
    rendering with the cache on must give the same frame as with it off (disjointness holds)."""
    W, H = 200, 150
    b = S.lion_scene(W, H, 0.45)
    objs, n, nbg, edges, points = b.arrays()
    for i in range(n):
        objs[i].id = i + 1
    a = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H), usecache=False)
    c = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H), usecache=True)
    assert np.array_equal(a, c)


def test_bbox_reject_is_neutral_on_benchmark_scenes(oracle):
    """render.ml:1270-1279 rejects on `bounds_of_basicshape`, which can be one pixel short of the shape
    (DESIGN.md "known divergences"); on the benchmark scenes it never changes a pixel."""
    W, H = 320, 240
    b = S.lion_scene(W, H, 0.7)
    objs, n, nbg, edges, points = b.arrays()
    a, ua = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H), bbox_reject=True, want_u=True)
    c, uc = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H), bbox_reject=False, want_u=True)
    assert np.array_equal(a, c) and np.array_equal(ua, uc)


def test_host_geometry_matches_oracle(oracle):
    """The product's host-side flattening / brush sampling (scene.py) against the oracle's restatement."""
    rng = random.Random(25)
    for _ in range(30):
        p = [(rng.uniform(0, 400), rng.uniform(0, 400)) for _ in range(4)]
        mine = []
        S.bezier_subdivide(0.2, p[0], p[1], p[2], p[3], mine)
        ref = oracle.flatten_bezier([c for q in p for c in q])
        assert np.array_equal(np.array([[a[0], a[1], b_[0], b_[1]] for a, b_ in mine]), ref)
        r = rng.choice([3.0, 4.5, 10.0, 20.0])
        segs = [("C", p[0], p[1], p[2], p[3]), ("L", p[3], p[0])]
        pts = S.brush_points(r, [segs])
        flat = [[1.0] + [c for q in p for c in q], [0.0, p[3][0], p[3][1], p[0][0], p[0][1], 0, 0, 0, 0]]
        ref = oracle.points_on_path(flat, float(int(np.ceil(r)) * 2 + 1) / 20.0)
        assert np.array_equal(np.array(pts, dtype=np.int32).reshape(-1, 2), ref)


def _dense_mask(flat, W, H):
    return util.bitmap_of_flat(flat, 0, 0, W, H).astype(bool)


def _dense_opacity(oracle, edges, winding, shape_flat, W, H):
    out = np.zeros((H, W), dtype=np.int64)
    vals = oracle.polygon_opacity(edges, winding, shape_flat)
    k = 0
    for y, spans in util.rows_of_flat(shape_flat):
        for x, l in spans:
            out[y, x:x + l] = vals[k:k + l]
            k += l
    return out


@pytest.mark.parametrize("op", ["union", "intersection", "subtraction", "xor"])
def test_cpg_sprite_equals_per_pixel_alpha_algebra(oracle, op):
    """The region-by-region restatement of sprite_of_cpg (render.ml:867-981) equals one per-pixel expression of the
    operands' mattes (a = 255 inside an operand's minshape, its antialiased opacity on the rest of its shape,
    0 outside) — the form the CUDA walker evaluates (cpg_alpha, raster_core.cuh)."""
    W, H = 200, 160
    a_sub = [S.polygon_segments([(10.2, 40.3), (190.6, 43.1), (188.0, 100.2), (12.0, 97.7)])]
    b_sub = [S.polygon_segments([(30.0, 20.4), (170.0, 70.2), (160.0, 140.9), (25.0, 72.6)])]
    col = S.rgba8(30, 30, 200)
    b = S.SceneBuilder()
    o = b.cpg(op, a_sub, b_sub, S.Fill.plain(col))
    objs, n, nbg, e, p = b.arrays()
    img = oracle.render_frame(objs, n, 0, e, p, (0, 0, W, H))
    ea, eb = e[o.first:o.first + o.count], e[o.first2:o.first2 + o.count2]
    sa, ma = oracle.shapeminshape(ea, 0)
    sb, mb = oracle.shapeminshape(eb, 0)
    SA, MA, SB, MB = (_dense_mask(f, W, H) for f in (sa, ma, sb, mb))
    A = np.where(MA, 255, np.where(SA, _dense_opacity(oracle, ea, 0, sa, W, H), 0))
    B = np.where(MB, 255, np.where(SB, _dense_opacity(oracle, eb, 0, sb, W, H), 0))
    if op == "union":
        shp, mn, al = SA | SB, MA | MB, np.minimum(255, A + B)
    elif op == "intersection":
        shp, mn, al = SA & SB, MA & MB, np.minimum(A, B)
    elif op == "subtraction":
        shp, mn, al = SA & ~MB, MA & ~SB, np.maximum(0, A - B)
    else:
        shp, mn = (SA | SB) & ~(MA & MB), (MB & ~SA) | (MA & ~SB)
        ia, ib = 255 - A, 255 - B
        al = np.where((A < 128) & (B < 128), np.maximum(A, B), np.where((A >= 128) & (B < 128), 255 - np.maximum(ia, B),
                      np.where((A < 128) & (B >= 128), 255 - np.maximum(A, ib), np.maximum(ia, ib))))
    al = np.where(mn, 255, al)
    exp = np.zeros((H, W), dtype=np.uint32)
    for y, x in np.argwhere(shp):
        exp[y, x] = oracle.colour_op("dissolve", col, int(al[y, x]))
    assert np.array_equal(img, exp)


def test_filter_limits(oracle):
    """Filters (render.ml:1080-1131): a hole with an opaque matte shows nothing of the scene below inside the
    geometry's opaque part; monochrome leaves grey pixels grey; a filter over an empty scene is transparent;
    the filter's whole shape leaves u."""
    import math

    W, H = 160, 120
    circ = [S.polygon_segments([(80.3 + 40.5 * math.cos(2 * math.pi * i / 24), 60.2 + 40.5 * math.sin(2 * math.pi * i / 24)) for i in range(24)])]

    def frame(kind, below, **kw):
        b = S.SceneBuilder()
        f = b.filter(kind, circ, **kw)
        for pts, c in below:
            b.polygon(pts, S.Fill.plain(c))
        objs, n, nbg, e, p = b.arrays()
        img, u = oracle.render_frame(objs, n, 0, e, p, (0, 0, W, H), want_u=True)
        shp, _ = oracle.shapeminshape(e[f.first:f.first + f.count], 0)
        return img, _dense_mask(u, W, H), _dense_mask(shp, W, H)

    quad = [(10.0, 10.0), (150.0, 12.0), (148.0, 110.0), (12.0, 108.0)]
    red, grey = S.rgba8(200, 30, 30), S.rgba8(90, 90, 90)
    img, u, g = frame("hole", [(quad, red)])
    assert not (u & g).any(), "the extra finish removes the whole geometry from u"
    assert img[60, 80] == 0 and img[60, 20] == red
    img_m, _, _ = frame("monochrome", [(quad, grey)])
    img_p, _, _ = frame("hole", [])
    assert img_m[60, 80] == grey and img_m[60, 20] == grey, "monochrome of grey is the identity"
    assert not img_p.any()
    img_b, _, _ = frame("blur", [(quad, red)], kernel=("gaussian", 3))
    assert img_b[60, 80] == red, "the blur of a flat colour is that colour"


def test_benchmark_scene_builders_render(oracle):
    """The C4 / C5 scene builders produce scenes the oracle accepts at a reduced size: filters with
    their reading-scene group, a Convolved page shadow, the lion group as the dragged object."""
    W, H = 320, 180
    objs, n, nbg, e, p = S.filter_scene(W, H, 0.55).arrays()
    kinds = [o.kind for o in objs]
    assert kinds.count(6) == 3 and any(o.filter_kind == 100 for o in objs) and any(o.convolve for o in objs)
    img = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H))
    assert (img >> 24 == 255).all(), "the background makes every pixel opaque"
    b, mover = S.drag_scene(W, H, 0.3, n_static=20)
    objs, n, nbg, e, p = b.arrays()
    assert objs[mover].kind == 2 and objs[mover].id == 1
    base = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H))
    k, depth = mover + 1, 1
    while depth:
        depth += 1 if objs[k].kind == 2 else -1 if objs[k].kind == 3 else 0
        if objs[k].kind not in (2, 3):
            objs[k].dx, objs[k].dy = 7, -4
        k += 1
    moved = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H))
    assert not np.array_equal(base, moved)


def test_smear_points_product_vs_oracle():
    """coh_host_smear_points (host geometry of the product) against the oracle's restatement of
    Brush.points_of_brushstroke_smear / find_smear_directions (brush.ml:239-283): pieces at most 2 apart, start points
    truncated, consecutive duplicates dropped."""
    from coherence_renderer_b200 import abi
    from oracle import pyoracle

    paths = [
        [("C", (40.0, 150.0), (90.0, 30.0), (150.0, 170.0), (200.0, 50.0))],
        [("L", (30.5, 40.25), (200.0, 60.0)), ("L", (200.0, 60.0), (120.75, 160.0)), ("C", (120.75, 160.0), (10.0, 10.0), (300.0, 5.0), (12.0, 90.0))],
        [("L", (5.0, 5.0), (6.0, 5.5))],
    ]
    for segs in paths:
        got = abi.host_smear_points(segs)
        ref = pyoracle.smear_points(abi._seg_records(segs))
        assert np.array_equal(got, ref)
        assert len(got) >= 1
        d = np.abs(np.diff(got.astype(int), axis=0))
        assert len(d) == 0 or (d.max() <= 2 and d.sum(axis=1).min() >= 1)   # adjacent pixels, no duplicates
