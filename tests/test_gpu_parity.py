"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle.
Bit-exact for span sets and integer coverage; RGBA within 1 LSB (north_star), and in fact
asserted bit-exact for plain fills."""
import random

import numpy as np
import pytest

from coherence_renderer_b200 import abi, scene as S
from tests import util

pytestmark = pytest.mark.gpu


def _max_lsb(a, b):
    return int(np.abs(a.view(np.uint8).astype(int) - b.view(np.uint8).astype(int)).max())


def test_shapeminshape_random_polygons(ctx, oracle):
    rng = random.Random(11)
    for it in range(150):
        edges = util.random_polygon_edges(rng)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        hs, hm = ctx.shapeminshape_of_edgelist(edges, w)
        got_s, got_m = ctx.shape_export(hs), ctx.shape_export(hm)
        assert np.array_equal(got_s, ref_s), f"shape differs (case {it})"
        assert np.array_equal(got_m, ref_m), f"minshape differs (case {it})"
        assert ctx.shape_card(hs) == oracle.shape_card(ref_s)
        ctx.shape_free(hs)
        ctx.shape_free(hm)


def _comb_edges(x0, y0, teeth, pitch, width, height):
    """A comb as ONE compound path: `teeth` thin rectangles `width` wide every `pitch` pixels (a line of small text has
    this crossing density)."""
    segs = []
    for k in range(teeth):
        xa = x0 + k * pitch
        segs.append(S.polygon_segments([(xa, y0), (xa + width, y0 + 0.3), (xa + width + 0.4, y0 + height), (xa + 0.2, y0 + height - 0.2)]))
    return np.concatenate([abi.host_edgelist_of_subpath(sg) for sg in segs]).reshape(-1, 4)


def test_many_crossings_per_window(ctx, oracle):
    """Dense compound paths: ~40 band crossings touch one 32-pixel window of a row (the crossing lists hold 64; the
    stand-alone scan falls back from its 256-pixel windows to single words) — stand-alone scan conversion, the matte,
    and whole frames through the three-phase path and the fused walker."""
    edges = _comb_edges(10.3, 8.1, 150, 1.55, 0.8, 22.0)    # ~20 teeth = 40 crossings per 32 pixels, 165 per 256
    for w in (0, 1):
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        hs, hm = ctx.shapeminshape_of_edgelist(edges, w)
        assert np.array_equal(ctx.shape_export(hs), ref_s) and np.array_equal(ctx.shape_export(hm), ref_m)
        ctx.shape_free(hs)
        ctx.shape_free(hm)
        maxshape = oracle.shape_op("difference", ref_s, ref_m)
        hx = ctx.shape_import(maxshape)
        assert np.array_equal(ctx.polygon_opacity(edges, w, hx), oracle.polygon_opacity(edges, w, maxshape))
        ctx.shape_free(hx)
    W, H = 300, 48
    for fused in (0, 1):
        ctx.set_option("fused", fused)
        try:
            b = S.SceneBuilder()
            b.path_edges(edges, S.Fill.plain(S.dissolve(S.rgba8(20, 20, 20), 230)))
            b.path_edges(_comb_edges(5.0, 20.0, 120, 2.1, 1.0, 20.0), S.Fill.plain(S.rgba8(200, 40, 40)), winding=S.COH_EVENODD)
            b.begin_background()
            b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
            got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
            assert np.array_equal(got_u, ref_u), fused
            assert _max_lsb(got, ref) == 0, fused
        finally:
            ctx.set_option("fused", -1)
    # beyond the lists' capacity the frame still fails loudly (never a wrong picture)
    dense = _comb_edges(10.0, 4.0, 400, 0.4, 0.2, 12.0)     # ~160 crossings per 32 pixels
    b = S.SceneBuilder()
    b.path_edges(dense, S.Fill.plain(S.rgba8(0, 0, 0)))
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    with pytest.raises(abi.CohError):
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.fb_read_rgba(0, 0, W, H)
    ctx.scene_free(sc)


def test_polygon_opacity_random(ctx, oracle):
    rng = random.Random(12)
    for it in range(40):
        edges = util.random_polygon_edges(rng, rmax=80.0)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        if len(ref_s) == 0:
            continue
        maxshape = oracle.shape_op("difference", ref_s, ref_m)
        hs = ctx.shape_import(maxshape)
        got = ctx.polygon_opacity(edges, w, hs)
        ref = oracle.polygon_opacity(edges, w, maxshape)
        assert np.array_equal(got, ref), f"AA opacity differs (case {it})"
        ctx.shape_free(hs)


def test_shape_algebra(ctx, oracle):
    rng = random.Random(13)
    for it in range(25):
        a, b = util.random_shape_flat(rng), util.random_shape_flat(rng, x0=20, y0=-10)
        ha, hb = ctx.shape_import(a), ctx.shape_import(b)
        assert np.array_equal(ctx.shape_export(ha), a)
        for name, fn in (("union", ctx.shape_union), ("difference", ctx.shape_difference), ("intersection", ctx.shape_intersection)):
            h = fn(ha, hb)
            assert np.array_equal(ctx.shape_export(h), oracle.shape_op(name, a, b)), name
            ctx.shape_free(h)
        m, n = rng.randint(0, 6), rng.randint(0, 6)
        h = ctx.shape_bloat(ha, m, n)
        assert np.array_equal(ctx.shape_export(h), oracle.shape_unary("bloat", a, m, n))
        ctx.shape_free(h)
        h = ctx.shape_erode(ha, m, n)
        assert np.array_equal(ctx.shape_export(h), oracle.shape_unary("erode", a, m, n))
        ctx.shape_free(h)
        h = ctx.shape_translate(ha, 7, -3)
        assert np.array_equal(ctx.shape_export(h), oracle.shape_unary("translate", a, 7, -3))
        ctx.shape_free(h)
        ctx.shape_free(ha)
        ctx.shape_free(hb)
    assert ctx.shape_box(0, 0, 0, 0) == 0
    with pytest.raises(abi.CohError):
        ctx.shape_box(0, 0, -1, 3)


def _render_both(ctx, oracle, b, W, H, update=None):
    update = update or (0, 0, W, H)
    objs, n, nbg, edges, points = b.arrays()
    ref, ref_u = oracle.render_frame(objs, n - nbg, nbg, edges, points, update, want_u=True)
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, update, abi.COH_RENDER_RECORD_U)
    ctx.sync()
    got = ctx.fb_read_rgba(update[0], update[1], update[2], update[3])
    hu = ctx.render_uncovered()
    got_u = ctx.shape_export(hu)
    ctx.shape_free(hu)
    ctx.scene_free(sc)
    return got, ref, got_u, ref_u


def test_lion_c1_frame(ctx, oracle):
    """C1: the lion at the prototype's 1280x1024 canvas."""
    W, H = 1280, 1024
    got, ref, got_u, ref_u = _render_both(ctx, oracle, S.lion_scene(W, H, 3.0), W, H)
    assert np.array_equal(got_u, ref_u), "covered-so-far span set differs"
    assert _max_lsb(got, ref) == 0


def test_lion_small_update_box(ctx, oracle):
    W, H = 640, 480
    got, ref, got_u, ref_u = _render_both(ctx, oracle, S.lion_scene(W, H, 1.4), W, H, update=(201, 77, 263, 301))
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_random_layered_polygons(ctx, oracle):
    """C3-style: translucent and opaque random polygons, front to back."""
    W, H = 512, 384
    for seed in (1, 2, 3):
        b = S.random_scene(W, H, 120, seed=seed, brush_fraction=0.0)
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u), f"u differs (seed {seed})"
        assert _max_lsb(got, ref) == 0, f"RGBA differs (seed {seed})"


def test_groups_and_pretrans(ctx, oracle):
    W, H = 256, 200
    b = S.SceneBuilder()
    b.polygon([(20.3, 20.1), (120.7, 30.2), (60.2, 150.9)], S.Fill.plain(S.dissolve(S.rgba8(200, 30, 30), 180)))
    b.group_begin(pretrans=140)
    b.polygon([(40.0, 40.0), (200.5, 60.5), (90.0, 180.0)], S.Fill.plain(S.rgba8(10, 200, 40)))
    b.polygon([(10.0, 100.0), (240.0, 90.0), (200.0, 190.0), (30.0, 170.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 0, 255), 100)), pretrans=200)
    b.group_begin()
    b.rectangle(S.rgba8(255, 255, 0), 100.0, 20.0, 180.0, 120.0, pretrans=90)
    b.group_end()
    b.group_end()
    b.polygon([(0.0, 0.0), (255.0, 0.0), (255.0, 199.0), (0.0, 199.0)], S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 60)))
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_fancy_fills(ctx, oracle):
    W, H = 200, 160
    b = S.SceneBuilder()
    grad = S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 0, 255), 128))
    rad = S.Fill.radial((100.0, 80.0), (100.0, 90.0), (160.0, 80.0), True, True, S.rgba8(255, 255, 0), S.rgba8(0, 80, 0))
    b.polygon([(30.5, 30.5), (170.2, 40.1), (150.0, 140.0), (40.0, 120.0)], grad)
    b.polygon([(10.0, 10.0), (190.0, 12.0), (180.0, 150.0), (15.0, 140.0)], rad)
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    # AA pixels of fancy fills sample the fill at their span start (polygon.ml:736) even when the span
    # began several 32-pixel tiles to the left: the walker's cross-tile carry makes this exact.
    assert _max_lsb(got, ref) == 0


@pytest.fixture
def options(ctx):
    """Force code paths through coh_set_option; everything is reset afterwards."""
    yield ctx.set_option
    for name, v in (("walk_h", 0), ("fused", -1), ("aa_general", 0), ("bin_cache", 1), ("comp_rows", 1), ("fork_prefill", 1)):
        ctx.set_option(name, v)


def test_fancy_fills_three_phase_path(ctx, oracle, options):
    """The three-phase frame path (forced) with gradient fills: the compositing walk keeps the cross-tile carry of
    span starts."""
    options("fused", 0)
    test_fancy_fills(ctx, oracle)
    test_fancy_fill_long_shallow_edges(ctx, oracle)


def test_fancy_fill_long_shallow_edges(ctx, oracle):
    """Edge runs hundreds of pixels long (nearly horizontal edges) under a gradient whose alpha varies."""
    W, H = 700, 120
    b = S.SceneBuilder()
    g1 = S.Fill.gradient((0.0, 0.0), (700.0, 0.0), True, True, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 255, 0), 90))
    g2 = S.Fill.gradient((50.0, 10.0), (650.0, 100.0), False, False, S.dissolve(S.rgba8(0, 0, 255), 200), S.rgba8(255, 255, 0))
    b.polygon([(5.2, 20.3), (690.7, 23.9), (680.1, 70.2), (15.5, 66.6)], g1)
    b.polygon([(20.0, 10.0), (660.0, 14.5), (600.0, 110.0), (100.0, 104.0)], g2, pretrans=200)
    b.polygon([(2.0, 50.0), (698.0, 52.0), (698.0, 58.0), (2.0, 57.0)], S.Fill.plain(S.rgba8(0, 0, 0)))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_brush_strokes(ctx, oracle):
    W, H = 320, 240
    b = S.SceneBuilder()
    b.brush(0.8, 6.0, [[("C", (30.0, 200.0), (100.0, 10.0), (220.0, 230.0), (290.0, 40.0))]], S.Fill.plain(S.rgba8(20, 20, 120)))
    b.brush(1.0, 3.0, [[("L", (10.0, 10.0), (300.0, 60.0))]], S.Fill.plain(S.dissolve(S.rgba8(200, 0, 0), 200)))
    b.polygon([(60.0, 60.0), (260.0, 80.0), (160.0, 220.0)], S.Fill.plain(S.rgba8(240, 200, 60)))
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_path_flattening_on_device(ctx, oracle):
    """N2: Polygon.edgelist_of_path (polygon.ml:83-127, 262-287; coord.ml:47) on the device — de Casteljau subdivision
    in FP64, one rounding per operation — gives the edges of the host-side geometry, in the same order; scan conversion
    of the resident edges gives the oracle's shape and minshape."""
    import random

    rnd = random.Random(77)
    for trial in range(12):
        segs, p0 = [], (rnd.uniform(20, 600), rnd.uniform(20, 400))
        start = p0
        for k in range(rnd.randint(2, 9)):
            if rnd.random() < 0.6:
                c = [(rnd.uniform(-50, 700), rnd.uniform(-50, 500)) for _ in range(3)]
                segs.append(("C", p0, c[0], c[1], c[2]))
                p0 = c[2]
            else:
                p1 = (rnd.uniform(0, 640), rnd.uniform(0, 480))
                segs.append(("L", p0, p1))
                p0 = p1
        segs.append(("L", p0, start))
        host = abi.host_edgelist_of_subpath(segs)
        dev = ctx.edgelist_of_path(segs)
        assert np.array_equal(dev, host), trial
        for w in (abi.COH_NONZERO, abi.COH_EVENODD):
            hs, hm = ctx.shapeminshape_of_path(segs, w)
            rs, rm = oracle.shapeminshape(host, w)
            assert np.array_equal(ctx.shape_export(hs), rs) and np.array_equal(ctx.shape_export(hm), rm), (trial, w)
            ctx.shape_free(hs)
            ctx.shape_free(hm)
    # degenerate curves (coincident control points: distances that are not FP_normal count as flat)
    flat = [("C", (10.0, 10.0), (10.0, 10.0), (10.0, 10.0), (10.0, 10.0)), ("C", (10.0, 10.0), (50.0, 10.0), (90.0, 10.0), (130.0, 10.0))]
    assert np.array_equal(ctx.edgelist_of_path(flat), abi.host_edgelist_of_subpath(flat))


def _brush_obj(opacity, radius, fill, kind=0):
    o = abi.CohObject()
    o.kind = abi.COH_OBJ_BRUSH
    o.winding = kind
    o.brush_opacity, o.brush_radius = float(opacity), float(radius)
    o.pretrans = -1
    o.id = -1
    fill.apply(o)
    return o


def test_brush_entry_points(ctx, oracle):
    """brush.mli:20-27 outside a scene: Brush.shape_of_brushstroke, sprite_of_brushstroke (plain and gradient fills, in
    the stroke's own shape and in another one) and Brush.smear of a sprite, against the oracle."""
    segs = [("C", (40.0, 150.0), (90.0, 30.0), (150.0, 170.0), (200.0, 50.0)), ("L", (200.0, 50.0), (230.0, 120.0))]
    for opacity, radius in ((0.8, 6.0), (1.0, 3.5)):
        pts = abi.host_brush_points(segs, radius)
        for fill in (S.Fill.plain(S.dissolve(S.rgba8(20, 200, 120), 220)), S.Fill.gradient((30.0, 30.0), (220.0, 160.0), True, True, S.rgba8(255, 0, 0), S.rgba8(0, 0, 255))):
            bo = _brush_obj(opacity, radius, fill)
            hs = ctx.brush_shape(bo, pts)
            ref_s = oracle.brush_shape(bo, pts)
            assert np.array_equal(ctx.shape_export(hs), ref_s)
            got = ctx.brush_sprite(bo, pts, hs)
            assert np.array_equal(got, oracle.brush_sprite(bo, pts, ref_s))
            box = ctx.shape_box(60, 40, 120, 90)
            part = ctx.shape_intersection(hs, box)
            assert np.array_equal(ctx.brush_sprite(bo, pts, part), oracle.brush_sprite(bo, pts, ctx.shape_export(part)))
            for h in (hs, box, part):
                ctx.shape_free(h)
    # dummy brush: white all over its shape
    bo = _brush_obj(1.0, 6.0, S.Fill.plain(S.rgba8(1, 2, 3)), kind=abi.COH_BRUSH_DUMMY)
    pts = abi.host_brush_points(segs, 6.0)
    hs = ctx.brush_shape(bo, pts)
    assert np.array_equal(ctx.shape_export(hs), oracle.brush_shape(bo, pts))
    got = ctx.brush_sprite(bo, pts, hs)
    assert np.array_equal(got, oracle.brush_sprite(bo, pts, ctx.shape_export(hs))) and (got == 0xFFFFFFFF).all()
    ctx.shape_free(hs)
    # Brush.smear of a sprite: a gradient-filled polygon's sprite, smeared along the stroke
    sm = abi.host_smear_points(segs)
    poly = abi.host_edgelist_of_subpath(S.polygon_segments([(30.3, 30.2), (210.5, 43.9), (188.1, 150.7), (28.8, 134.4)]))
    ps, pm = ctx.shapeminshape_of_edgelist(poly, abi.COH_NONZERO)
    fo = _brush_obj(1.0, 1.0, S.Fill.gradient((30.0, 30.0), (220.0, 160.0), True, True, S.rgba8(255, 200, 0), S.dissolve(S.rgba8(0, 0, 255), 150)))
    spr = ctx.polygon_sprite(fo, poly, abi.COH_NONZERO, ps)
    for opacity, radius in ((1.0, 7.0), (0.6, 3.0)):
        bo = _brush_obj(opacity, radius, S.Fill.plain(S.WHITE))
        pts = abi.host_brush_points(segs, radius)
        ho, got = ctx.brush_smear(ps, spr, bo, pts, sm)
        ref_shape, ref = oracle.brush_smear(ctx.shape_export(ps), spr, bo, pts, sm)
        assert np.array_equal(ctx.shape_export(ho), ref_shape)
        assert np.array_equal(got, ref)
        ctx.shape_free(ho)
    # NullSprite: only the stroke's shape, clear
    ho, got = ctx.brush_smear(0, np.zeros(0, np.uint32), bo, pts, sm)
    ref_shape, ref = oracle.brush_smear(np.zeros(0, np.int32), np.zeros(0, np.uint32), bo, pts, sm)
    assert np.array_equal(ctx.shape_export(ho), ref_shape) and np.array_equal(got, ref) and not got.any()
    for h in (ho, ps, pm):
        ctx.shape_free(h)


def test_dummy_brush(ctx, oracle):
    """Brushstroke with a Dummy brush (brush.ml:14-22, 70-73, 178-181): the whole shape of the stroke — the boxes around
    its stamp points — in opaque white, whatever fill it was given; minshape null."""
    W, H = 320, 240
    b = S.SceneBuilder()
    b.brush(0.8, 6.0, [[("C", (30.0, 200.0), (100.0, 10.0), (220.0, 230.0), (290.0, 40.0))]], S.Fill.plain(S.rgba8(20, 20, 120)))
    b.dummy_brush(4.5, [[("C", (20.0, 30.0), (120.0, 220.0), (200.0, 20.0), (300.0, 200.0))]], pretrans=170)
    b.dummy_brush(9.0, [[("L", (10.0, 10.0), (300.0, 60.0))]])
    b.polygon([(60.0, 60.0), (260.0, 80.0), (160.0, 220.0)], S.Fill.plain(S.rgba8(240, 200, 60)))
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0
    assert (got == 0xFFFFFFFF).sum() > 2000   # the opaque white stroke shows


def test_rgb888_export(ctx):
    W, H = 96, 64
    b = S.SceneBuilder()
    b.rectangle(S.rgba8(10, 20, 30), 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    rgb = ctx.fb_read_rgb888(3, 5, 40, 20)
    assert rgb.shape == (20, 40, 3) and (rgb == np.array([10, 20, 30], dtype=np.uint8)).all()
    ctx.scene_free(sc)


def test_band_partition_equals_whole_frame(ctx, oracle):
    """§8e: rendering N bands separately and concatenating equals the single-band frame."""
    W, H = 640, 480
    b = S.lion_scene(W, H, 1.4)
    objs, n, nbg, edges, points = b.arrays()
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.fb_configure(W, H)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    whole = ctx.fb_read_rgba(0, 0, W, H).copy()
    parts = []
    for k in range(3):
        y0, y1 = k * H // 3, (k + 1) * H // 3
        ctx.fb_configure(W, H, y0, y1)
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        parts.append(ctx.fb_read_rgba(0, y0, W, y1 - y0).copy())
    assert np.array_equal(np.concatenate(parts, axis=0), whole)
    ctx.scene_free(sc)


def test_band_partition_of_filter_scenes(ctx, oracle):
    """§8e for scenes with filter objects: a band renders its own rows; what its lenses read above and below the band
    (a blur's reading shape, a smear's) is rendered again on that band's context — halo rows are recomputed, not
    exchanged.  Bands cut through every lens; the assembled strips equal the whole frame and the oracle."""
    W, H = 200, 160
    b, _ = _filter_scene("blur", W, H, second=("monochrome", {}), kernel=("gaussian", 3))
    b.smear_filter(1.0, 5.0, [[("L", (20.0, 30.0), (180.0, 130.0))]])
    b.polygon([(15.0, 20.0), (190.0, 25.0), (180.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(90, 20, 160), 210)))
    _finish(b, W, H)
    objs, n, nbg, edges, points = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    whole = ctx.fb_read_rgba(0, 0, W, H).copy()
    assert _max_lsb(whole, ref) == 0
    for cuts in ([0, 53, 107, 160], [0, 80, 81, 97, 160]):
        parts = []
        for y0, y1 in zip(cuts[:-1], cuts[1:]):
            ctx.fb_configure(W, H, y0, y1)
            ctx.render_frame(sc, (0, 0, W, H))
            ctx.sync()
            parts.append(ctx.fb_read_rgba(0, y0, W, y1 - y0).copy())
        assert np.array_equal(np.concatenate(parts, axis=0), whole), cuts
    # a partial update inside a band
    ctx.fb_configure(W, H, 40, 120)
    ctx.render_frame(sc, (30, 20, 140, 130))
    ctx.sync()
    assert np.array_equal(ctx.fb_read_rgba(30, 40, 140, 80), whole[40:120, 30:170])
    ctx.fb_configure(W, H)
    ctx.scene_free(sc)


def test_pretrans_group_returns_pixels_to_u(ctx, oracle):
    """A PreTrans group with opaque members gives its pixels back to the parent's u when it closes;
    many objects follow so that several scan passes start while the group is open."""
    W, H = 160, 96
    b = S.SceneBuilder()
    b.group_begin(pretrans=120)
    for i in range(7):
        b.polygon([(10.0 + 9 * i, 8.0), (150.0 - 5 * i, 14.0 + 3 * i), (140.0, 88.0 - 4 * i), (12.0 + 3 * i, 80.0)], S.Fill.plain(S.rgba8(30 * i, 255 - 30 * i, 90)))
    b.group_end()
    for i in range(9):
        b.polygon([(5.0 + 11 * i, 5.0 + 2 * i), (60.0 + 11 * i, 9.0), (70.0 + 9 * i, 90.0), (8.0 + 10 * i, 70.0)], S.Fill.plain(S.dissolve(S.rgba8(200, 20 * i, 255 - 20 * i), 255 if i % 2 else 150)))
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_large_random_scene_object_parallel_binning(ctx, oracle):
    """> 1024 leaves: K1 switches to warp-per-leaf binning + per-cell sort; brush strokes use their
    per-row stamp ranges.  C3-style scene (polygons and brush strokes, translucent and opaque)."""
    W, H = 1024, 768
    b = S.random_scene(W, H, 1500, seed=7)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_translation_alias_equals_translated_edges(ctx, oracle):
    """Cache.addtranslation (cache.ml:423-436): an alias (dx, dy) serves the cached shape/sprite translated by
    whole pixels, i.e. exactly the original edges moved by 32*d sub-bins."""
    W, H = 400, 300
    pts = [(60.3, 40.2), (180.9, 70.1), (150.0, 200.7), (40.0, 160.0)]
    for dx, dy in ((37, 21), (-45, -30), (0, 64)):
        b = S.SceneBuilder()
        b.polygon(pts, S.Fill.plain(S.dissolve(S.rgba8(30, 90, 200), 200)), dx=dx, dy=dy)
        b.brush(0.9, 5.0, [[("C", (20.0, 250.0), (120.0, 20.0), (260.0, 280.0), (380.0, 60.0))]], S.Fill.plain(S.rgba8(200, 40, 40)), dx=dx, dy=dy)
        b.begin_background()
        b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
        # and, while every coordinate stays positive, the same picture as moving the integer edges themselves
        # by 32 sub-bins per pixel.  (Moving the FLOAT path is not equivalent — sub_of_float (f + d) may round
        # differently, SURVEY.md App. B — and neither is moving edges into negative coordinates, where
        # pix_of_sub and toint truncate toward zero: the alias keeps the ORIGINAL position's rounding.)
        if dx < 0 or dy < 0:
            continue
        poly_only = S.SceneBuilder()
        o = poly_only.polygon(pts, S.Fill.plain(S.dissolve(S.rgba8(30, 90, 200), 200)), dx=dx, dy=dy)
        objs, n, nbg, edges, points = poly_only.arrays()
        b2 = S.SceneBuilder()
        b2.path_edges(edges + np.array([32 * dx, 32 * dy, 32 * dx, 32 * dy], dtype=np.int32), S.Fill.plain(S.dissolve(S.rgba8(30, 90, 200), 200)))
        objs2, n2, nbg2, edges2, points2 = b2.arrays()
        ref2 = oracle.render_frame(objs2, n2 - nbg2, nbg2, edges2, points2, (0, 0, W, H))
        ctx.fb_configure(W, H)
        sc = ctx.scene_create(objs, nbg, edges, points)
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        assert np.array_equal(ctx.fb_read_rgba(0, 0, W, H), ref2)
        ctx.scene_free(sc)


def test_drag_sequence_dirty_regions_and_cache(ctx, oracle):
    """C4 in miniature: an object is dragged by integer pixels; every frame re-renders only
    the dirty region plaindirty = ((shp_o - minshp_n) ∪ (shp_n - minshp_o)) ∩ u computed from cached,
    HBM-resident span sets, and the framebuffer must equal a full render of the moved scene."""
    W, H = 480, 360
    b = S.random_scene(W, H, 60, seed=11, brush_fraction=0.0, background=False)
    mover = b.polygon([(100.2, 80.1), (220.5, 100.9), (200.0, 230.3), (90.0, 200.0)], S.Fill.plain(S.rgba8(250, 220, 30)), oid=7)
    mover_index = len(b.objs) - 1
    for i in range(25):  # objects behind the mover
        x, y = 30.0 + 16 * i, 20.0 + 11 * i
        b.polygon([(x, y), (x + 90.5, y + 10.2), (x + 70.1, y + 95.5), (x - 5.0, y + 60.0)], S.Fill.plain(S.dissolve(S.rgba8(20 * i % 255, 200, 255 - 9 * i), 255 if i % 3 else 170)), oid=100 + i)
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    ctx.cache_clear()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    master = ctx.shape_box(0, 0, W, H)
    mo = objs[mover_index]
    medges = edges[mo.first : mo.first + mo.count]
    tx = ty = 0
    for step, (dx, dy) in enumerate([(3, 2), (5, -1), (-7, 4), (0, 9), (12, 12), (-30, -20)]):
        so, mno = ctx.scene_object_shape(sc, mover_index)
        ctx.scene_translate_object(sc, mover_index, dx, dy)
        tx, ty = tx + dx, ty + dy
        sn, mnn = ctx.scene_object_shape(sc, mover_index)
        # the new shape is the cached one translated — equal to scan-converting the moved integer edges
        ref_sn, ref_mn = oracle.shapeminshape(medges + np.array([32 * tx, 32 * ty, 32 * tx, 32 * ty], dtype=np.int32), mo.winding)
        assert np.array_equal(ctx.shape_export(sn), ref_sn) and np.array_equal(ctx.shape_export(mnn), ref_mn)
        dirty = ctx.dirty_region(so, mno, sn, mnn, master, plain=True)
        ref_so, ref_mo = oracle.shapeminshape(medges + np.array([32 * (tx - dx), 32 * (ty - dy)] * 2, dtype=np.int32), mo.winding)
        box = util.flat_of_rows([(y, [(0, W)]) for y in range(H)])
        ref_dirty = oracle.shape_op("intersection", oracle.shape_op("union", oracle.shape_op("difference", ref_so, ref_mn), oracle.shape_op("difference", ref_sn, ref_mo)), box)
        assert np.array_equal(ctx.shape_export(dirty), ref_dirty), f"dirty region differs at step {step}"
        ctx.render_frame_shape(sc, dirty)
        ctx.sync()
        got = ctx.fb_read_rgba(0, 0, W, H)
        objs[mover_index].dx, objs[mover_index].dy = tx, ty
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert _max_lsb(got, ref) == 0, f"frame differs after incremental update {step}"
        for h in (so, mno, sn, mnn, dirty):
            ctx.shape_free(h)
    st = ctx.cache_stats()
    assert st["shape_hits"] >= 11 and st["shape_misses"] == 1 and st["entries"] == 1 and st["bytes"] > 0
    ctx.shape_free(master)
    ctx.scene_free(sc)


def test_update_shape_with_holes(ctx, oracle):
    """render_frame over a non-rectangular update (rows with several spans): only those pixels change."""
    W, H = 256, 192
    b = S.lion_scene(W, H, 0.55)
    objs, n, nbg, edges, points = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    blank = S.SceneBuilder()
    blank.rectangle(S.rgba8(1, 2, 3), 0.0, 0.0, float(W), float(H))
    bo, bn, bnbg, be, bp = blank.arrays()
    sc0 = ctx.scene_create(bo, bnbg, be, bp)
    ctx.render_frame(sc0, (0, 0, W, H))
    rows = [(y, [(10 + (y % 7), 40), (70, 3), (100 + y % 5, 90 - y % 11)]) for y in range(20, 170) if y % 9]
    upd = util.flat_of_rows(rows)
    hu = ctx.shape_import(upd)
    ctx.render_frame_shape(sc, hu)
    ctx.sync()
    got = ctx.fb_read_rgba(0, 0, W, H)
    full = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H), bbox_reject=False)
    mask = util.bitmap_of_flat(upd, 0, 0, W, H)
    assert np.array_equal(got[mask], full[mask])
    assert (got[~mask] == S.rgba8(1, 2, 3)).all()
    ctx.shape_free(hu)
    ctx.scene_free(sc)
    ctx.scene_free(sc0)


def test_convolved_objects(ctx, oracle):
    """Convolved (kernel, Basic (plain fill, Path)) — the page-shadow construct of engine.ml:85-88: shape =
    bloat r r, minshape = erode r r, sprite = AA raster of the bloated region, X pass, Y pass, cropped
    (render.ml:536-555, 1023-1052; convolve.ml:115-258)."""
    W, H = 320, 240
    for kern in (("gaussian", 4), ("unit", 3), ("gaussian", 7)):
        b = S.SceneBuilder()
        b.polygon([(120.3, 40.2), (250.9, 90.1), (150.0, 200.7)], S.Fill.plain(S.dissolve(S.rgba8(250, 180, 20), 220)))
        b.polygon([(50.0, 40.0), (200.0, 45.5), (190.5, 160.0), (60.0, 150.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 0, 0), 120)), convolve=kern, pretrans=230)
        b.polygon([(20.0, 100.0), (300.0, 110.0), (280.0, 230.0), (30.0, 220.0)], S.Fill.plain(S.rgba8(40, 160, 90)), convolve=("unit", 2), dx=-3, dy=5)
        b.begin_background()
        b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u), kern
        assert _max_lsb(got, ref) == 0, kern


def test_convolved_groups(ctx, oracle):
    """Convolved (kernel, Group members) (render.ml:63, 1023-1052 with a Group child): shape = bloat r r of the union of
    the members' shapes, minshape null (a Group's fill is "fancy", render.ml:536-555), sprite = convolve_sprite of the
    group's sprite.  Members: translucent and opaque paths, a primitive, a brush stroke, a CPG, a nested PreTrans group,
    a nested Convolved path; the Convolved group itself under PreTrans, translated, partly hidden and partly off-frame."""
    W, H = 360, 260
    for kern in (("gaussian", 3), ("unit", 2)):
        b = S.SceneBuilder()
        b.polygon([(200.3, 10.2), (340.9, 60.1), (250.0, 120.7)], S.Fill.plain(S.rgba8(250, 180, 20)))   # opaque, in front
        b.group_begin(convolve=kern, pretrans=240, oid=11)
        b.polygon([(60.0, 40.0), (230.0, 45.5), (210.5, 170.0), (70.0, 150.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 40, 200), 180)))
        b.rectangle(S.rgba8(200, 30, 30), 100.0, 90.0, 160.0, 140.0)
        b.group_begin(pretrans=128)
        b.polygon([(150.0, 60.0), (300.0, 80.0), (280.0, 200.0)], S.Fill.plain(S.rgba8(30, 160, 60)))
        b.polygon([(160.0, 100.0), (260.0, 110.0), (200.0, 190.0)], S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 90)), convolve=("unit", 2))
        b.group_end()
        b.brush(0.8, 6.0, [[("C", (80.0, 200.0), (140.0, 120.0), (220.0, 230.0), (300.0, 150.0))]], S.Fill.plain(S.rgba8(10, 10, 10)))
        b.group_end()
        b.group_begin(convolve=("gaussian", 5), dx=-40, dy=30)   # reaches out of the frame on the left / bottom
        b.polygon([(10.0, 150.0), (120.0, 160.0), (90.0, 250.0), (20.0, 240.0)], S.Fill.plain(S.rgba8(120, 0, 160)))
        b.polygon([(40.0, 170.0), (150.0, 200.0), (60.0, 230.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 200, 200), 200)))
        b.group_end()
        b.begin_background()
        b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u), kern
        assert _max_lsb(got, ref) == 0, kern
        # partial update, and the object's shape through the cache entry points
        objs, n, nbg, e, p = b.arrays()
        ctx.fb_configure(W, H)
        sc = ctx.scene_create(objs, nbg, e, p)
        ctx.render_frame(sc, (90, 70, 150, 120))
        sub = ctx.fb_read_rgba(90, 70, 150, 120)
        assert np.array_equal(sub, ref[70:190, 90:240]), kern
        ctx.scene_free(sc)
    # Render.shape_of_basicshape of the object: bloat r r (union of the members' shapes), minshape null
    b = S.SceneBuilder()
    b.group_begin(convolve=("gaussian", 5))
    b.polygon([(10.0, 150.0), (120.0, 160.0), (90.0, 250.0), (20.0, 240.0)], S.Fill.plain(S.rgba8(120, 0, 160)))
    b.polygon([(140.0, 170.0), (250.0, 200.0), (160.0, 230.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 200, 200), 200)))
    b.group_end()
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    gs, gm = ctx.scene_object_shape(sc, 0)
    sa, _ = oracle.shapeminshape(e[objs[1].first:objs[1].first + objs[1].count], objs[1].winding)
    sb, _ = oracle.shapeminshape(e[objs[2].first:objs[2].first + objs[2].count], objs[2].winding)
    assert np.array_equal(ctx.shape_export(gs), oracle.shape_unary("bloat", oracle.shape_op("union", sa, sb), 5, 5)) and gm == 0
    ctx.shape_free(gs)
    ctx.scene_free(sc)
    # what is refused, loudly
    b = S.SceneBuilder()
    b.group_begin(convolve=("gaussian", 3))
    b.polygon([(30.5, 30.5), (170.2, 40.1), (150.0, 140.0)], S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.rgba8(0, 0, 255)))
    b.group_end()
    objs, n, nbg, e, p = b.arrays()
    with pytest.raises(abi.CohError):
        ctx.scene_create(objs, nbg, e, p)


def test_convolve_sprite(ctx, oracle):
    """Convolve.convolve_sprite on arbitrary sprites (random canonical shapes, random premultiplied colours)."""
    rng = random.Random(41)
    for kern, r in (("gaussian", 3), ("unit", 2), ("gaussian", 5), ("unit", 6)):
        shp = util.random_shape_flat(rng, x0=-10, y0=5, w=140, h=70, density=0.5)
        n = oracle.shape_card(shp)
        a = np.array([rng.randint(0, 255) for _ in range(n)], dtype=np.uint32)
        rgba = np.array([(rng.randint(0, int(v)) | (rng.randint(0, int(v)) << 8) | (rng.randint(0, int(v)) << 16) | (int(v) << 24)) for v in a], dtype=np.uint32)
        ref_shape, ref_px = oracle.convolve_sprite(kern, r, shp, rgba)
        h = ctx.shape_import(shp)
        ho, got_px = ctx.convolve_sprite(kern, r, h, rgba)
        assert np.array_equal(ctx.shape_export(ho), ref_shape)
        assert np.array_equal(got_px, ref_px), (kern, r)
        ctx.shape_free(h)
        ctx.shape_free(ho)
    with pytest.raises(abi.CohError):
        ctx.convolve_sprite("unit", 0, 0, np.zeros(0, np.uint32))


def test_stroked_path_winding_quirk(ctx, oracle):
    """Basic (fill, StrokedPath ...): shape with NonZero, sprite with EvenOdd (render.ml:510 vs 1018).  The edge
    list here is a self-overlapping outline (what a stroker emits for a tight bend), where the rules differ."""
    W, H = 220, 180
    pts = [(20.0, 20.0), (200.0, 30.0), (190.0, 160.0), (30.0, 150.0), (20.0, 20.0), (110.0, 10.0), (180.0, 90.0), (100.0, 170.0), (25.0, 95.0)]
    s = S.sub_of_float
    edges = np.array([[s(pts[i][0]), s(pts[i][1]), s(pts[(i + 1) % len(pts)][0]), s(pts[(i + 1) % len(pts)][1])] for i in range(len(pts))], dtype=np.int32)
    b = S.SceneBuilder()
    b.stroked_path_edges(edges, S.Fill.plain(S.dissolve(S.rgba8(10, 10, 10), 210)))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    b2 = S.SceneBuilder()
    b2.path_edges(edges, S.Fill.plain(S.dissolve(S.rgba8(10, 10, 10), 210)))
    b2.begin_background()
    b2.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, e, p = b2.arrays()
    assert not np.array_equal(oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H)), ref), "the quirk must be visible on this input"


def _cpg_scene(op, W, H, fill, **kw):
    b = S.SceneBuilder()
    sq = [S.polygon_segments([(30.3, 30.2), (120.5, 33.9), (118.1, 110.7), (28.8, 104.4)])]
    star = [S.polygon_segments([(80.0, 20.0), (180.4, 80.2), (90.7, 150.1), (150.2, 25.5), (60.9, 120.3)])]
    b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
    b.cpg(op, sq, star, fill, winding_b=S.COH_EVENODD, **kw)
    b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    return b


@pytest.mark.parametrize("op", ["union", "intersection", "subtraction", "xor"])
def test_cpg_objects(ctx, oracle, op):
    """Basic (fill, CPG (op, Path a, Path b)) (render.ml:522-528 shapes, 867-981 sprite): overlapping operands
    with different winding rules, plain and gradient fills, under an alias offset and a dissolving group."""
    W, H = 200, 160
    for fill in (S.Fill.plain(S.dissolve(S.rgba8(200, 30, 30), 230)),
                 S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 0, 255), 128))):
        b = _cpg_scene(op, W, H, fill)
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u), op
        assert _max_lsb(got, ref) == 0, op
    b = _cpg_scene(op, W, H, S.Fill.plain(S.rgba8(200, 30, 30)), dx=13, dy=-7, pretrans=150)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, op
    # shallow edges: a minshape pixel of an operand need not be fully covered (the minshape's row band misses
    # the top quarter of the AA window); the reference never consults an operand's matte inside its minshape
    b = S.SceneBuilder()
    b.cpg(op, [S.polygon_segments([(10.2, 40.3), (190.6, 43.1), (188.0, 100.2), (12.0, 97.7)])],
          [S.polygon_segments([(30.0, 20.4), (170.0, 70.2), (160.0, 140.9), (25.0, 72.6)])], S.Fill.plain(S.rgba8(30, 30, 200)))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, op


def test_cpg_object_shape(ctx, oracle):
    """coh_scene_object_shape of a CPG object = the reference's set expressions of the operands' shapes."""
    W, H = 200, 160
    ctx.fb_configure(W, H)
    for op in ("union", "intersection", "subtraction", "xor"):
        b = _cpg_scene(op, W, H, S.Fill.plain(S.rgba8(1, 2, 3)))
        objs, n, nbg, e, p = b.arrays()
        sc = ctx.scene_create(objs, nbg, e, p)
        hs, hm = ctx.scene_object_shape(sc, 1)
        o = objs[1]
        sa, ma = oracle.shapeminshape(e[o.first:o.first + o.count], o.winding)
        sb, mb = oracle.shapeminshape(e[o.first2:o.first2 + o.count2], o.winding2)
        U, D, I = (lambda x, y: oracle.shape_op("union", x, y)), (lambda x, y: oracle.shape_op("difference", x, y)), (lambda x, y: oracle.shape_op("intersection", x, y))
        exp_s, exp_m = {"union": (U(sa, sb), U(ma, mb)), "intersection": (I(sa, sb), I(ma, mb)), "subtraction": (D(sa, mb), D(ma, sb)),
                        "xor": (D(U(sa, sb), I(ma, mb)), U(D(mb, sa), D(ma, sb)))}[op]
        assert np.array_equal(ctx.shape_export(hs), exp_s), op
        assert np.array_equal(ctx.shape_export(hm), exp_m), op
        ctx.shape_free(hs)
        ctx.shape_free(hm)
        ctx.scene_free(sc)


def _circle(cx, cy, r, n=28):
    import math

    return [S.polygon_segments([(cx + r * math.cos(2 * math.pi * i / n), cy + r * math.sin(2 * math.pi * i / n)) for i in range(n)])]


def _filter_scene(kind, W, H, second=None, matte=None, **kw):
    b = S.SceneBuilder()
    b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
    f = b.filter(kind, _circle(100.3, 80.2, 50.5), fill=matte, **kw)
    b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
    if second is not None:
        b.filter(second[0], _circle(120.0, 100.0, 40.0), **second[1])
    b.group_begin(pretrans=200)
    b.polygon([(60.0, 20.0), (120.0, 140.0), (20.0, 120.0)], S.Fill.plain(S.rgba8(250, 240, 20)))
    b.rectangle(S.rgba8(0, 0, 0), 90.0, 100.0, 170.0, 150.0)
    b.group_end()
    b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
    return b, f


def _finish(b, W, H):
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    return b


@pytest.mark.parametrize("kind,kw", [("hole", {}), ("monochrome", {}), ("blur", {"kernel": ("gaussian", 3)}), ("blur", {"kernel": ("unit", 2)})])
def test_filters(ctx, oracle, kind, kw):
    """Filter objects (render.ml:1080-1131, blend' 1248-1265; filters.ml hole / monochrome / blur): the scene below
    is rendered twice (reading scene and ordinary scene), filtered and blended by the geometry's AA matte; the
    filter's whole shape leaves u.  Opaque and translucent mattes."""
    W, H = 200, 160
    for matte in (None, S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 170))):
        b, _ = _filter_scene(kind, W, H, matte=matte, **kw)
        got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
        assert np.array_equal(got_u, ref_u), (kind, kw)
        assert _max_lsb(got, ref) == 0, (kind, kw)


def test_filters_stacked_and_partial_update(ctx, oracle):
    """A filter below another filter is rendered inside both of the upper filter's recursive renders."""
    W, H = 200, 160
    b, _ = _filter_scene("blur", W, H, second=("monochrome", {}), kernel=("gaussian", 2))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    b, _ = _filter_scene("monochrome", W, H, second=("hole", {}))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H, update=(70, 50, 60, 70))
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0


def test_filter_with_reading_scene(ctx, oracle):
    """Filters whose reading scene is a rewrite of the objects below (affine, rgb, wireframe, swapdepth, minus:
    filters.ml:105-212, 271-332): the caller supplies the rewritten list as a reading-scene group."""
    W, H = 200, 160
    b, f = _filter_scene("scene", W, H)
    b.reading_scene_begin(f)   # "affine": the objects below, moved and squashed
    b.polygon([(40.3, 50.2), (160.5, 53.9), (158.1, 100.7), (38.8, 94.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
    b.polygon([(15.0, 25.0), (195.0, 28.0), (185.0, 120.0), (22.0, 110.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
    b.group_end()
    got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0


def test_filters_inside_groups(ctx, oracle):
    """A filter object that is a member of a Group (render.ml:988-1001): its objects below are the rest of the group's
    own list; the group (PreTrans or not, nested or not) is composited as one sprite."""
    W, H = 200, 160
    for kind, kw, pt in (("monochrome", {}, -1), ("blur", {"kernel": ("gaussian", 2)}, 180), ("hole", {}, -1)):
        b = S.SceneBuilder()
        b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
        b.group_begin(pretrans=pt)
        b.polygon([(50.3, 20.2), (90.5, 23.9), (88.1, 60.7), (48.8, 54.4)], S.Fill.plain(S.rgba8(30, 30, 220)))
        b.filter(kind, _circle(100.3, 80.2, 40.5), **kw)
        b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
        b.group_begin()
        b.filter("monochrome", _circle(70.0, 100.0, 25.0))
        b.polygon([(60.0, 20.0), (120.0, 140.0), (20.0, 120.0)], S.Fill.plain(S.dissolve(S.rgba8(250, 240, 20), 200)))
        b.group_end()
        b.group_end()
        b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
        assert np.array_equal(got_u, ref_u), kind
        assert _max_lsb(got, ref) == 0, kind
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H, update=(40, 30, 90, 100))
        assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, kind


def test_filter_minus(ctx, oracle):
    """Filters.minus (filters.ml:289-303), a single-object hole: inside shape (filter) ∩ shape (object below) the
    scene is read without that object; the rest of the filter's shape is finished with nothing in it.  Head object a
    path, a group, and an object the filter does not meet."""
    W, H = 200, 160
    for variant in range(3):
        b = S.SceneBuilder()
        b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
        b.filter("minus", _circle(100.3, 80.2, 45.5), fill=S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 255 if variant != 1 else 150)))
        if variant == 0:
            b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
        elif variant == 1:
            b.group_begin(pretrans=220)
            b.polygon([(30.3, 30.2), (120.5, 33.9), (118.1, 100.7), (28.8, 94.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
            b.polygon([(80.0, 70.0), (170.0, 75.0), (160.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(30, 30, 200), 180)))
            b.group_end()
        else:
            b.polygon([(2.0, 2.0), (30.0, 3.0), (20.0, 25.0)], S.Fill.plain(S.rgba8(200, 30, 30)))
        b.polygon([(60.0, 20.0), (120.0, 140.0), (20.0, 120.0)], S.Fill.plain(S.rgba8(250, 240, 20)))
        b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
        assert np.array_equal(got_u, ref_u), variant
        assert _max_lsb(got, ref) == 0, variant


def test_smear_filter(ctx, oracle):
    """Filters.smear (filters.ml:201-217) = Brush.smear (brush.ml:235-331) along a brush stroke over the scene below:
    the sequential read-block / blend-block walk over the smear points, twice, on the reference's canvas (steps whose
    blocks leave it are skipped, as its swallowed exceptions do)."""
    W, H = 240, 200
    strokes = [
        (1.0, 7.0, [[("C", (40.0, 150.0), (90.0, 30.0), (150.0, 170.0), (200.0, 50.0))]]),
        (0.7, 4.0, [[("L", (30.0, 40.0), (200.0, 60.0)), ("L", (200.0, 60.0), (120.0, 160.0))]]),
        (1.0, 3.0, [[("L", (205.5, 12.0), (232.0, 17.5)), ("C", (232.0, 17.5), (236.0, 40.0), (200.0, 30.0), (215.0, 60.0))]]),   # mostly over nothing
    ]
    for opacity, radius, path in strokes:
        b = S.SceneBuilder()
        b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
        b.smear_filter(opacity, radius, path)
        b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
        b.polygon([(100.0, 20.0), (220.0, 150.0), (60.0, 180.0)], S.Fill.plain(S.dissolve(S.rgba8(250, 240, 20), 200)))
        b.rectangle(S.rgba8(0, 0, 0), 150.0, 100.0, 230.0, 190.0)
        if radius == 4.0:   # a scene that is not flat: the compositing walker, fused and three-phase
            b.group_begin(pretrans=200)
            b.polygon([(20.0, 100.0), (120.0, 110.0), (60.0, 190.0)], S.Fill.plain(S.rgba8(10, 200, 90)))
            b.polygon([(40.0, 20.0), (230.0, 40.0), (200.0, 90.0)], S.Fill.plain(S.dissolve(S.rgba8(90, 10, 200), 140)))
            b.group_end()
        got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
        assert np.array_equal(got_u, ref_u), radius
        assert _max_lsb(got, ref) == 0, radius
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H, update=(60, 40, 120, 90))
        assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, radius
        if radius == 7.0:
            first = b
    b = first
    # the filter does something: without it the frame differs
    b2 = S.SceneBuilder()
    b2.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
    b2.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
    b2.polygon([(100.0, 20.0), (220.0, 150.0), (60.0, 180.0)], S.Fill.plain(S.dissolve(S.rgba8(250, 240, 20), 200)))
    b2.rectangle(S.rgba8(0, 0, 0), 150.0, 100.0, 230.0, 190.0)
    plain, _, _, _ = _render_both(ctx, oracle, _finish(b2, W, H), W, H)
    full, _, _, _ = _render_both(ctx, oracle, b, W, H)
    assert (plain != full).sum() > 200


def test_filters_with_object_geometry(ctx, oracle):
    """Filter geometry other than a path (COH_GEOM_NEXT): a brush stroke (examples.ml:279-300 monobrush: the matte is the
    stroke's Gaussian alpha), a Convolved path (engine.ml:33-36: a soft-edged lens), a CPG and a group — the object that
    follows the filter is its geometry, its sprite's alpha the matte (render.ml:1099)."""
    W, H = 200, 160

    def scene(kind, geometry, kw):
        b = S.SceneBuilder()
        b.polygon([(10.0, 60.0), (190.0, 70.0), (100.0, 95.0)], S.Fill.plain(S.dissolve(S.rgba8(0, 90, 200), 120)))
        f = b.filter_with_geometry(kind, **kw)
        geometry(b)
        b.polygon([(30.3, 30.2), (150.5, 33.9), (148.1, 130.7), (28.8, 124.4)], S.Fill.plain(S.rgba8(200, 30, 30)))
        b.polygon([(60.0, 20.0), (120.0, 140.0), (20.0, 120.0)], S.Fill.plain(S.dissolve(S.rgba8(250, 240, 20), 200)))
        b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
        return _finish(b, W, H), f

    white = S.Fill.plain(S.WHITE)
    geometries = {
        "brush": lambda b: b.brush(0.9, 9.0, [[("C", (30.0, 130.0), (80.0, 20.0), (130.0, 150.0), (180.0, 40.0))]], white),
        "convolved": lambda b: b.path(_circle(100.3, 80.2, 40.5), white, abi.COH_NONZERO, convolve=("gaussian", 4)),
        "cpg": lambda b: (b.group_begin(), b.cpg("xor", _circle(80.0, 80.0, 40.0), _circle(120.0, 85.0, 35.0), S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 200))), b.group_end()),
        "path": lambda b: b.path(_circle(100.3, 80.2, 40.5), S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 190)), abi.COH_NONZERO),
        "group": lambda b: (b.group_begin(), b.polygon([(40.0, 40.0), (90.0, 45.0), (85.0, 110.0)], white),
                            b.brush(1.0, 5.0, [[("L", (100.0, 30.0), (170.0, 120.0))]], white), b.group_end()),
    }
    for gname, geometry in geometries.items():
        for kind, kw in (("monochrome", {}), ("blur", {"kernel": ("gaussian", 2)}), ("minus", {})):
            b, f = scene(kind, geometry, kw)
            got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
            assert np.array_equal(got_u, ref_u), (gname, kind)
            assert _max_lsb(got, ref) == 0, (gname, kind)
    # translated, and a partial update
    b, f = scene("monochrome", geometries["brush"], {})
    f.dx, f.dy = -14, 9
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H, update=(40, 30, 120, 100))
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0


def test_translated_and_dragged_lenses(ctx, oracle):
    """A filter object with an alias offset (render.ml:259-271) reads its geometry moved by whole pixels; dragging a
    lens (coh_scene_drag_object: alldirty of its shape at both places) re-renders exactly what a full frame of the
    moved scene shows."""
    W, H = 200, 160
    for kind, kw in (("monochrome", {}), ("blur", {"kernel": ("gaussian", 2)})):
        b, f = _filter_scene(kind, W, H, **kw)
        f.dx, f.dy = 23, -11
        got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
        assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, kind
    b, f = _filter_scene("monochrome", W, H)
    _finish(b, W, H)
    objs, n, nbg, edges, points = b.arrays()
    fi = [k for k, o in enumerate(objs) if o.kind == abi.COH_OBJ_FILTER][0]
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    for step, (dx, dy) in enumerate([(5, 3), (-12, 7), (30, -20), (1, 0)]):
        bb = ctx.scene_drag_object(sc, fi, dx, dy)
        ctx.sync()
        assert bb[2] >= bb[0] and bb[3] >= bb[1]
        objs[fi].dx += dx
        objs[fi].dy += dy
        got = ctx.fb_read_rgba(0, 0, W, H)
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert _max_lsb(got, ref) == 0, f"frame differs after lens drag step {step}"
    # shapeonly_of_basicshape of the filter = its geometry's shape, where the alias has taken it
    hs, hm = ctx.scene_object_shape(sc, fi)
    g0, g1 = ctx.shapeminshape_of_edgelist(np.asarray(edges).reshape(-1, 4)[objs[fi].first:objs[fi].first + objs[fi].count], objs[fi].winding)
    gt = ctx.shape_translate(g0, objs[fi].dx, objs[fi].dy)
    assert np.array_equal(ctx.shape_export(hs), ctx.shape_export(gt))
    for h in (hs, hm, g0, g1, gt):
        ctx.shape_free(h)
    ctx.scene_free(sc)


def test_drag_object_fused_step(ctx, oracle):
    """coh_scene_drag_object = translate + dirty region + render of that region, on device-resident span sets: after
    every step the framebuffer equals the oracle's full render of the moved scene — for a plain path (plaindirty),
    a group (alldirty) and a primitive.  (Plain fills only: an antialiased pixel of a fancy fill takes the fill at
    the start of its visible span, polygon.ml:736, so in the reference too a partial update can differ from a
    full render.)"""
    W, H = 400, 300
    b = S.SceneBuilder()
    b.polygon([(100.2, 80.1), (220.5, 100.9), (200.0, 230.3), (90.0, 200.0)], S.Fill.plain(S.dissolve(S.rgba8(250, 220, 30), 200)), oid=7)
    b.group_begin(oid=8)
    b.polygon([(40.0, 40.0), (140.5, 60.5), (90.0, 160.0)], S.Fill.plain(S.rgba8(10, 200, 40)))
    b.polygon([(60.0, 90.0), (180.0, 95.0), (150.0, 190.0)], S.Fill.plain(S.dissolve(S.rgba8(255, 0, 0), 140)))
    b.group_end()
    b.rectangle(S.rgba8(0, 0, 90), 250.0, 40.0, 330.0, 120.0, oid=9)
    for i in range(12):
        x, y = 20.0 + 28 * i, 30.0 + 17 * i
        b.polygon([(x, y), (x + 90.5, y + 10.2), (x + 70.1, y + 95.5), (x - 5.0, y + 60.0)], S.Fill.plain(S.dissolve(S.rgba8(20 * i % 255, 200, 255 - 9 * i), 255 if i % 3 else 170)), oid=100 + i)
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    group_index, prim_index = 1, 5
    ctx.cache_clear()
    ctx.cache_configure(True, 64 << 20)
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    moves = [(0, 3, 2), (group_index, -4, 5), (prim_index, 6, -3), (0, -9, 0), (group_index, 11, 7), (prim_index, -20, 15), (0, 1, 1)]
    for step, (idx, dx, dy) in enumerate(moves):
        bb = ctx.scene_drag_object(sc, idx, dx, dy)
        ctx.sync()
        assert bb[2] >= bb[0] and bb[3] >= bb[1]
        last = idx
        if objs[idx].kind == abi.COH_OBJ_GROUP_BEGIN:  # the members carry the offsets
            k, depth = idx + 1, 1
            while depth:
                if objs[k].kind == abi.COH_OBJ_GROUP_BEGIN:
                    depth += 1
                elif objs[k].kind == abi.COH_OBJ_GROUP_END:
                    depth -= 1
                else:
                    objs[k].dx += dx
                    objs[k].dy += dy
                k += 1
        else:
            objs[last].dx += dx
            objs[last].dy += dy
        got = ctx.fb_read_rgba(0, 0, W, H)
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert _max_lsb(got, ref) == 0, f"frame differs after drag step {step}"
    ctx.scene_free(sc)
    ctx.cache_clear()


def test_dirty_filter(ctx, oracle):
    """Render.dirty_filter (render.ml:1418-1438): the dirty region of a moved object grows through every blur lens
    in front of it (bloatdirty, filters.ml:63-75) and passes the other filters unchanged."""
    W, H = 200, 160
    b, _ = _filter_scene("blur", W, H, second=("monochrome", {}), kernel=("gaussian", 3))
    objs, n, nbg, e, p = _finish(b, W, H).arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    rng = random.Random(5)
    f = objs[1]
    fshape, _ = oracle.shapeminshape(e[f.first:f.first + f.count], f.winding)
    r = 3
    for it in range(6):
        d = util.random_shape_flat(rng, x0=10 + 20 * it, y0=5 + 12 * it, w=90, h=70, density=0.5)
        hd = ctx.shape_import(d)
        # lmo = the last scene object: both filters are in front of it
        got = ctx.dirty_filter(sc, n - nbg - 1, hd)
        bf = oracle.shape_unary("bloat", fshape, r, r)
        inf, outf = oracle.shape_op("intersection", bf, d), oracle.shape_op("difference", d, bf)
        exp = oracle.shape_op("union", oracle.shape_op("intersection", oracle.shape_unary("bloat", inf, r, r), bf), outf)
        assert np.array_equal(ctx.shape_export(got), exp), it
        ctx.shape_free(got)
        # lmo in front of every filter: nothing applies
        got = ctx.dirty_filter(sc, 0, hd)
        assert np.array_equal(ctx.shape_export(got), d)
        ctx.shape_free(got)
        ctx.shape_free(hd)
    ctx.scene_free(sc)


def test_async_readback_equals_blocking_read(ctx):
    """coh_fb_read_rgba_async / coh_fb_read_wait: frames read back while the next one renders are the frames
    that were rendered (two staging buffers; three reads in flight exercise the reuse path)."""
    W, H = 320, 200
    ctx.fb_configure(W, H)
    frames, scenes = [], []
    for k in range(3):
        b = S.lion_scene(W, H, 0.6 + 0.2 * k)
        objs, n, nbg, e, p = b.arrays()
        scenes.append(ctx.scene_create(objs, nbg, e, p))
    outs = [np.zeros((H, W), dtype=np.uint32) for _ in range(3)]
    for k in range(3):
        ctx.render_frame(scenes[k], (0, 0, W, H))
        ctx.fb_read_rgba_async(0, 0, W, H, outs[k])
    ctx.fb_read_wait()
    for k in range(3):
        ctx.render_frame(scenes[k], (0, 0, W, H))
        assert np.array_equal(ctx.fb_read_rgba(0, 0, W, H), outs[k]), k
        ctx.scene_free(scenes[k])


@pytest.mark.parametrize("walk_h,fused,aa_general,comp_rows", [("1", "1", 0, 1), ("4", "1", 0, 1), ("16", "1", 0, 1), ("4", "0", 0, 1), ("4", "0", 1, 1), ("4", "0", 0, 0)])
def test_every_walker_variant(ctx, oracle, walk_h, fused, aa_general, comp_rows, options):
    """The walker is compiled for work items of 1, 4 and 16 rows, and plain polygon scenes have a three-phase path
    (scan / visibility / antialiasing kernels + a compositing walk) next to the fused one; the library picks by
    scene and frame size.  Every variant must give the same pixels (COH_WALK_H / COH_FUSED force one)."""
    options("walk_h", int(walk_h))
    options("fused", int(fused))
    options("comp_rows", comp_rows)     # flat scenes in three-phase frames: row compositor, or the walker
    options("aa_general", aa_general)   # antialiasing of the three-phase frame: interval form + general kernel, or all general
    W, H = 640, 480
    b = S.lion_scene(W, H, 1.4, pretrans=200)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    # flat scenes (the lion's group is the first member of the list and dissolves into it; a plain list of
    # translucent polygons with a PreTrans member)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, S.lion_scene(W, H, 1.4), W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    b = S.random_scene(300, 200, 60, seed=9, brush_fraction=0.0)
    b.objs[3].pretrans = 120
    b.objs[7].pretrans = 0
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, 300, 200)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, 300, 200, update=(37, 21, 150, 101))
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    b = S.random_scene(300, 200, 50, seed=5, brush_fraction=0.3)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, 300, 200)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    W, H = 200, 160
    b = S.SceneBuilder()
    b.polygon([(30.5, 30.5), (170.2, 40.1), (150.0, 140.0), (40.0, 120.0)], S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 0, 255), 128)))
    b.cpg("xor", [S.polygon_segments([(10.2, 40.3), (190.6, 43.1), (188.0, 100.2), (12.0, 97.7)])],
          [S.polygon_segments([(30.0, 20.4), (170.0, 70.2), (160.0, 140.9), (25.0, 72.6)])], S.Fill.plain(S.rgba8(30, 30, 200)))
    b.filter("blur", _circle(100.3, 80.2, 50.5), kernel=("gaussian", 2))
    b.polygon([(5.0, 5.0), (195.0, 8.0), (185.0, 150.0), (12.0, 140.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 160, 20), 90)))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0


@pytest.mark.parametrize("fused", [-1, 0, 1])
def test_peer_framebuffer_mirror(ctx, oracle, options, fused):
    """coh_fb_set_peers: every pixel rendered into the framebuffer is also stored to the peer framebuffers (on a
    multi-GPU box: the other ranks' frames over NVLink; here a second buffer on the same GPU stands in for a
    peer).  Two bands rendered by two passes into mirrored buffers assemble the whole frame in both — through the
    fused walker and through the three-phase frame with the row compositor (whose blocks then also finish the
    background cells: k_prefill is off with peers)."""
    import torch

    options("fused", fused)

    W, H = 320, 208
    b = S.lion_scene(W, H, 0.75)
    objs, n, nbg, e, p = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H), bbox_reject=False)
    bufs = [torch.zeros((H, W), dtype=torch.int32, device="cuda") for _ in range(2)]
    sc = None
    try:
        for k, (y0, y1) in enumerate([(0, 112), (112, H)]):  # "rank" k renders band k into buffer k, mirrored to the other
            ctx.fb_configure(W, H, y0, y1)
            ctx.fb_attach(bufs[k].data_ptr())
            ctx.fb_set_peers([bufs[1 - k].data_ptr()])
            if sc is None:
                sc = ctx.scene_create(objs, nbg, e, p)
            ctx.render_frame(sc, (0, 0, W, H))
            ctx.sync()
        for k in range(2):
            got = bufs[k].cpu().numpy().view(np.uint32)
            assert _max_lsb(got, ref) == 0, f"buffer {k}"
    finally:
        ctx.fb_set_peers([])
        ctx.fb_configure(W, H)
        ctx.fb_attach(0)         # back to a context-owned framebuffer for the tests that follow
        if sc:
            ctx.scene_free(sc)


def test_peer_framebuffer_mirror_of_filter_frames(ctx, oracle):
    """Bands of a scene with filter objects, gathered through peer framebuffers: the filter kernels do not mirror their
    stores, every band sends its finished rows to the peers in one strip copy."""
    import torch

    W, H = 200, 160
    b, _ = _filter_scene("blur", W, H, second=("monochrome", {}), kernel=("gaussian", 2))
    _finish(b, W, H)
    objs, n, nbg, e, p = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H))
    bufs = [torch.zeros((H, W), dtype=torch.int32, device="cuda") for _ in range(2)]
    sc = None
    try:
        for k, (y0, y1) in enumerate([(0, 90), (90, H)]):
            ctx.fb_configure(W, H, y0, y1)
            ctx.fb_attach(bufs[k].data_ptr())
            ctx.fb_set_peers([bufs[1 - k].data_ptr()])
            if sc is None:
                sc = ctx.scene_create(objs, nbg, e, p)
            ctx.render_frame(sc, (0, 0, W, H))
            ctx.sync()
        for k in range(2):
            got = bufs[k].cpu().numpy().view(np.uint32)
            assert _max_lsb(got, ref) == 0, f"buffer {k}"
    finally:
        ctx.fb_set_peers([])
        ctx.fb_configure(W, H)
        ctx.fb_attach(0)
        if sc:
            ctx.scene_free(sc)


def test_lion_c2_full_size_frame(ctx, oracle):
    """BASELINE.json's headline configuration at full size: the lion at 3840x2160 (scale 7), whole-frame update —
    framebuffer and covered-so-far set bit-exact against the oracle."""
    W, H = 3840, 2160
    b = S.lion_scene(W, H, 7.0)
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u)
    assert _max_lsb(got, ref) == 0


def test_edge_cases(ctx, oracle):
    """Empty scenes, empty and out-of-frame updates, objects outside the frame or with no area, horizontal and
    zero-length edges, a frame narrower than one tile, group nesting up to and beyond the walker's accumulator stack."""
    W, H = 70, 37   # not multiples of the 32 x 16 cell
    b = S.SceneBuilder()
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)          # nothing at all
    assert not got.any() and np.array_equal(got_u, ref_u)
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)          # background only
    assert np.array_equal(got, ref) and np.array_equal(got_u, ref_u)
    b = S.SceneBuilder()
    b.polygon([(-300.0, -300.0), (-200.0, -310.0), (-250.0, -200.0)], S.Fill.plain(S.rgba8(255, 0, 0)))   # outside
    b.polygon([(10.0, 10.0), (60.0, 10.0), (35.0, 10.0)], S.Fill.plain(S.rgba8(0, 255, 0)))                 # no area
    b.path_edges(np.array([[320, 320, 320, 320], [320, 640, 1600, 640]], dtype=np.int32), S.Fill.plain(S.rgba8(0, 0, 255)))  # degenerate
    b.polygon([(-20.5, 5.2), (90.0, 8.0), (30.0, 60.0)], S.Fill.plain(S.dissolve(S.rgba8(200, 100, 0), 200)))  # crosses every border
    for _ in range(5):
        b.group_begin(pretrans=250)
    b.polygon([(5.0, 3.0), (66.0, 4.0), (40.0, 33.0)], S.Fill.plain(S.rgba8(1, 2, 3)))
    for _ in range(5):
        b.group_end()
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H, update=(33, 7, 20, 11))   # a box inside
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    ctx.render_frame(sc, (0, 0, 0, 0))             # Sprite.box x y 0 0 = NullShape: nothing happens
    ctx.render_frame(sc, (500, 500, 10, 10))       # entirely outside the frame
    ctx.sync()
    with pytest.raises(abi.CohError):
        ctx.render_frame(sc, (0, 0, -1, 5))        # Sprite.box: negative argument (sprite.ml:463)
    ctx.scene_free(sc)
    tri = [(5.0, 3.0), (66.0, 4.0), (40.0, 33.0)]
    deep = S.SceneBuilder()
    for _ in range(9):                             # PreTrans groups keep their accumulators; beyond the walker's stack
        deep.group_begin(pretrans=200)             # (MAX_DEPTH) the inner groups become canvases rendered once —
    deep.polygon(tri, S.Fill.plain(S.rgba8(1, 2, 3)))   # the reference has no nesting limit
    deep.polygon([(t[0] - 3.0, t[1] + 2.5) for t in tri], S.Fill.plain(S.dissolve(S.rgba8(90, 20, 200), 120)))
    for _ in range(9):
        deep.group_end()
    deep.polygon([(t[0] + 4.0, t[1] - 1.5) for t in tri], S.Fill.plain(S.dissolve(S.rgba8(20, 90, 30), 77)))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, deep, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    deep = S.SceneBuilder()
    for _ in range(12):                            # first members composited with plain Over dissolve into their parent:
        deep.group_begin()                         # any depth of those is fine
    deep.polygon(tri, S.Fill.plain(S.dissolve(S.rgba8(10, 200, 30), 100)))
    for _ in range(12):
        deep.group_end()
    deep.polygon([(t[0] + 9.5, t[1] + 4.25) for t in tri], S.Fill.plain(S.dissolve(S.rgba8(200, 20, 30), 180)))
    got, ref, got_u, ref_u = _render_both(ctx, oracle, deep, W, H)
    assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0
    # size limits of the packed (tile, row) words and of the 32-bit tap sums fail loudly
    with pytest.raises(abi.CohError):
        ctx.fb_configure(64, 70000)
    ctx.fb_configure(W, H)
    with pytest.raises(abi.CohError):
        ctx.convolve_sprite("gaussian", 65, ctx.shape_box(0, 0, 4, 4), np.zeros(16, np.uint32))


def test_filter_matte_interior_shortcut(ctx, oracle):
    """The matte of a filter's geometry is super-sampled only where an edge can be near (outside the twice-eroded
    minshape); geometries with acute spikes, thin slivers, holes (even-odd) and parts outside the frame must
    still match the reference, which samples every pixel."""
    W, H = 260, 200
    spike = [S.polygon_segments([(20.0, 20.0), (240.0, 35.0), (130.0, 60.0), (250.0, 180.0), (120.0, 90.0), (15.0, 190.0), (90.0, 70.0)])]
    ring = [S.polygon_segments([(-30.0, -20.0), (200.0, -10.0), (210.0, 170.0), (-25.0, 160.0)]),
            S.polygon_segments([(40.0, 40.0), (150.0, 45.0), (145.0, 120.0), (45.0, 118.0)])]
    sliver = [S.polygon_segments([(10.0, 100.0), (250.0, 101.2), (250.0, 106.9), (10.0, 104.1)])]
    for geom, wind in ((spike, S.COH_NONZERO), (ring, S.COH_EVENODD), (sliver, S.COH_NONZERO)):
        for kind, kw in (("monochrome", {}), ("blur", {"kernel": ("unit", 2)})):
            b = S.SceneBuilder()
            b.filter(kind, geom, fill=S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 240)), winding=wind, **kw)
            b.polygon([(5.0, 5.0), (255.0, 8.0), (245.0, 190.0), (12.0, 180.0)], S.Fill.plain(S.rgba8(200, 60, 20)))
            b.polygon([(60.0, 0.0), (200.0, 100.0), (60.0, 199.0)], S.Fill.plain(S.dissolve(S.rgba8(20, 60, 220), 150)))
            got, ref, got_u, ref_u = _render_both(ctx, oracle, _finish(b, W, H), W, H)
            assert np.array_equal(got_u, ref_u) and _max_lsb(got, ref) == 0, (kind, wind)


def test_no_device_memory_leaks(ctx, oracle):
    """Scenes, frames (plain, filtered, dragged), span-set algebra, cache entries and convolutions give back every
    byte they allocate: device memory in use returns to its level after repeated cycles."""
    W, H = 320, 240
    rng = random.Random(3)

    def cycle():
        ctx.cache_clear()
        ctx.cache_configure(True, 32 << 20)
        ctx.fb_configure(W, H)
        b, _ = _filter_scene("blur", W, H, second=("monochrome", {}), kernel=("gaussian", 2))
        objs, n, nbg, e, p = _finish(b, W, H).arrays()
        sc = ctx.scene_create(objs, nbg, e, p)
        ctx.render_frame(sc, (0, 0, W, H), abi.COH_RENDER_RECORD_U)
        hu = ctx.render_uncovered()
        ctx.shape_free(hu)
        ctx.scene_free(sc)
        b = S.random_scene(W, H, 40, seed=rng.randint(1, 99), brush_fraction=0.3)
        objs, n, nbg, e, p = b.arrays()
        for o in objs:
            if o.kind in (abi.COH_OBJ_PATH, abi.COH_OBJ_BRUSH):
                o.id = 500 + rng.randint(0, 10000)
        sc = ctx.scene_create(objs, nbg, e, p)
        ctx.render_frame(sc, (0, 0, W, H))
        for k in range(3):
            ctx.scene_drag_object(sc, k, 2, -1)
        hs, hm = ctx.scene_object_shape(sc, 0)
        hb = ctx.shape_bloat(hs, 2, 2) if hs else 0
        hd = ctx.shape_difference(hb, hm) if hb else 0
        for h in (hs, hm, hb, hd):
            ctx.shape_free(h)
        ctx.scene_free(sc)
        ctx.cache_clear()
        ctx.sync()

    for _ in range(3):
        cycle()
    base = ctx.mem_in_use()
    for _ in range(10):
        cycle()
    assert ctx.mem_in_use() == base


def test_object_shapes_of_brush_and_convolved(ctx, oracle):
    """Render.shape_of_basicshape for Brushstroke (dilated stamp centres, minshape null, brush.ml:135-173) and
    Convolved (bloat r r shape, erode r r minshape, render.ml:536-555)."""
    W, H = 300, 200
    b = S.SceneBuilder()
    b.brush(0.8, 6.0, [[("C", (30.0, 40.0), (120.0, 10.0), (200.0, 190.0), (270.0, 60.0))]], S.Fill.plain(S.rgba8(10, 10, 200)))
    quad = [(60.3, 50.2), (220.5, 63.9), (208.1, 150.7), (48.8, 134.4)]
    b.polygon(quad, S.Fill.plain(S.rgba8(200, 30, 30)), convolve=("gaussian", 4))
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    hs, hm = ctx.scene_object_shape(sc, 0)
    o = objs[0]
    pts = p[o.first:o.first + o.count]
    r = int(np.ceil(6.0))
    rows = {}
    for x, y in pts:
        for yy in range(y - r, y + r + 1):
            rows.setdefault(yy, []).append((x - r, 2 * r + 1))
    exp = None
    for yy in sorted(rows):
        for x, l in rows[yy]:
            one = util.flat_of_rows([(yy, [(x, l)])])
            exp = one if exp is None else oracle.shape_op("union", exp, one)
    assert np.array_equal(ctx.shape_export(hs), exp) and hm == 0
    ctx.shape_free(hs)
    hs, hm = ctx.scene_object_shape(sc, 1)
    o = objs[1]
    ss, mm = oracle.shapeminshape(e[o.first:o.first + o.count], o.winding)
    assert np.array_equal(ctx.shape_export(hs), oracle.shape_unary("bloat", ss, 4, 4))
    assert np.array_equal(ctx.shape_export(hm), oracle.shape_unary("erode", mm, 4, 4))
    ctx.shape_free(hs)
    ctx.shape_free(hm)
    ctx.scene_free(sc)


def test_binning_kept_with_the_scene_is_invalidated(ctx, oracle, options):
    """Whole-frame cell binning is kept with the scene (a pure function of the object boxes and the frame geometry).
    It must be rebuilt when an object moves or the framebuffer geometry changes, and switching it off must give the
    same pixels."""
    W, H = 640, 480
    b = S.lion_scene(W, H, 1.4)
    b.objs[0].id = 5
    objs, n, nbg, edges, points = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    try:
        for rep in range(3):  # the second and third frame run from the kept lists
            ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert _max_lsb(ctx.fb_read_rgba(0, 0, W, H), ref) == 0
        ctx.scene_translate_object(sc, 0, 37, -21)   # the whole lion group
        for k in range(1, n - nbg - 1):
            objs[k].dx, objs[k].dy = 37, -21
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
        assert _max_lsb(ctx.fb_read_rgba(0, 0, W, H), ref) == 0, "stale cell lists after a move"
        ctx.fb_configure(W, H, 100, 300)             # another band of the same frame
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        assert _max_lsb(ctx.fb_read_rgba(0, 100, W, 200), ref[100:300]) == 0, "stale cell lists after a change of band"
        ctx.fb_configure(W, H)
        options("bin_cache", 0)
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        assert _max_lsb(ctx.fb_read_rgba(0, 0, W, H), ref) == 0
    finally:
        ctx.set_option("bin_cache", 1)
        ctx.scene_free(sc)


def test_first_member_groups_dissolve_into_their_parent(ctx, oracle):
    """A Group that is the first member of its list and is composited with plain Over lands on a clear accumulator
    (`over clear s = s`), so the library lets its members composite straight into the parent's accumulator.  Translucent
    fills make any regrouping of `over` visible: nested first members, a later sibling group (kept), a first member
    with PreTrans (kept), and a group that is first only after a null object."""
    W, H = 260, 200

    def tri(b, x, y, c, a, **kw):
        return b.polygon([(x, y), (x + 110.5, y + 14.2), (x + 40.1, y + 120.7)], S.Fill.plain(S.dissolve(c, a)), **kw)

    for variant in range(3):
        b = S.SceneBuilder()
        if variant == 2:
            b.path_edges(np.zeros((0, 4), np.int32), S.Fill.plain(S.WHITE))   # NullShape object in front
        b.group_begin(pretrans=(150 if variant == 1 else None))
        b.group_begin()
        tri(b, 20.3, 15.1, S.rgba8(220, 40, 40), 130)
        tri(b, 50.9, 30.4, S.rgba8(40, 220, 40), 170)
        b.group_end()
        tri(b, 80.2, 22.8, S.rgba8(40, 40, 220), 90)
        b.group_end()
        b.group_begin()
        tri(b, 60.6, 50.5, S.rgba8(200, 200, 30), 140)
        tri(b, 100.1, 40.9, S.rgba8(30, 200, 200), 200, pretrans=99)
        b.group_end()
        tri(b, 10.0, 60.0, S.rgba8(255, 255, 255), 255)
        b.begin_background()
        b.rectangle(S.dissolve(S.rgba8(90, 90, 90), 200), 0.0, 0.0, float(W), float(H))
        b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
        got, ref, got_u, ref_u = _render_both(ctx, oracle, b, W, H)
        assert np.array_equal(got_u, ref_u), f"variant {variant}"
        assert _max_lsb(got, ref) == 0, f"variant {variant}"


def test_polygon_sprite_entry_point(ctx, oracle):
    """coh_polygon_sprite = Polygon.polygon_sprite_edgelist (polygon.ml:729-746): RGBA8 per pixel of the given shape in
    span order — plain, axial and radial fills (the fill is taken at every span's first x, polygon.ml:736)."""
    rng = random.Random(77)
    fills = [S.Fill.plain(S.dissolve(S.rgba8(200, 40, 90), 170)),
             S.Fill.gradient((10.0, 20.0), (150.0, 90.0), True, False, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 0, 255), 128)),
             S.Fill.radial((80.0, 60.0), (90.0, 60.0), (150.0, 60.0), True, True, S.rgba8(255, 255, 0), S.rgba8(0, 90, 30))]
    checked = 0
    for it in range(24):
        edges = util.random_polygon_edges(rng, lo=-10, hi=180, rmax=70, kmax=8)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        if len(ref_s) == 0:
            continue
        o = abi.CohObject()
        fills[it % 3].apply(o)
        for shp_flat in (ref_s, oracle.shape_op("difference", ref_s, ref_m)):   # the whole shape; only its max-shape
            if len(shp_flat) == 0:
                continue
            h = ctx.shape_import(shp_flat)
            got = ctx.polygon_sprite(o, edges, w, h)
            ref = oracle.polygon_sprite(o, edges, w, shp_flat)
            assert np.array_equal(got, ref), f"sprite differs (case {it}, fill {it % 3})"
            checked += len(ref)
            ctx.shape_free(h)
    assert checked > 10000


def test_rgb888_and_sprite_export_against_oracle(ctx, oracle):
    """coh_fb_read_rgb888 = what Wxgui.plot_sprite writes (wxgui.ml:417-424: the premultiplied r, g, b bytes of every
    pixel), coh_fb_read_sprite = the Sprite.sprite Render.render_frame returns (pixels of the update shape in span
    order) — both against the oracle's frame."""
    W, H = 320, 240
    b = S.lion_scene(W, H, 0.7)
    b.polygon([(30.5, 20.2), (290.1, 60.7), (120.9, 220.3)], S.Fill.plain(S.dissolve(S.rgba8(30, 60, 220), 120)))
    objs, n, nbg, edges, points = b.arrays()
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    x, y, w, h = 37, 21, 201, 150
    rgb = ctx.fb_read_rgb888(x, y, w, h)
    want = ref[y : y + h, x : x + w]
    assert np.array_equal(rgb[:, :, 0], (want & 255).astype(np.uint8)) and np.array_equal(rgb[:, :, 1], ((want >> 8) & 255).astype(np.uint8))
    assert np.array_equal(rgb[:, :, 2], ((want >> 16) & 255).astype(np.uint8))
    # a sprite on an update shape with several spans per row
    rng = random.Random(5)
    flat = util.random_shape_flat(rng, x0=5, y0=8, w=300, h=220)
    hs = ctx.shape_import(flat)
    got = ctx.fb_read_sprite(hs)
    want = np.concatenate([ref[yy, xx : xx + l] for yy, spans in util.rows_of_flat(flat) for xx, l in spans])
    assert np.array_equal(got, want)
    assert len(ctx.fb_read_sprite(0)) == 0
    outside = ctx.shape_box(W - 5, 3, 10, 4)
    with pytest.raises(abi.CohError):
        ctx.fb_read_sprite(outside)
    for s_ in (hs, outside):
        ctx.shape_free(s_)
    ctx.scene_free(sc)


def test_cache_entry_points(ctx, oracle):
    """Cache.addshape / getshape / addtranslation through the ABI (cache.ml:280-324, 370-388, 423-436): copies are kept,
    an alias serves the target's shapes translated, aliases of aliases collapse, statistics count."""
    rng = random.Random(3)
    ctx.cache_clear()
    ctx.cache_configure(True, 8 << 20)
    a, m = util.random_shape_flat(rng), util.random_shape_flat(rng, density=0.15)
    ha, hm = ctx.shape_import(a), ctx.shape_import(m)
    st0 = ctx.cache_stats()                   # the counters run for the life of the context (cache.ml:24-38)
    assert ctx.cache_getshape(41) is None
    ctx.cache_addshape(41, ha, hm)
    ctx.shape_free(ha)
    ctx.shape_free(hm)                       # the cache keeps copies of its own
    got = ctx.cache_getshape(41)
    assert got is not None
    assert np.array_equal(ctx.shape_export(got[0]), a) and np.array_equal(ctx.shape_export(got[1]), m)
    for h in got:
        ctx.shape_free(h)
    ctx.cache_addtranslation(42, 41, 7, -3)
    ctx.cache_addtranslation(43, 42, 10, 10)  # alias of an alias: (17, 7) from 41 (cache.ml:434-436)
    for oid, (dx, dy) in ((42, (7, -3)), (43, (17, 7))):
        got = ctx.cache_getshape(oid)
        assert got is not None
        assert np.array_equal(ctx.shape_export(got[0]), oracle.shape_unary("translate", a, dx, dy))
        assert np.array_equal(ctx.shape_export(got[1]), oracle.shape_unary("translate", m, dx, dy))
        for h in got:
            ctx.shape_free(h)
    st = ctx.cache_stats()
    assert st["shape_hits"] - st0["shape_hits"] == 3 and st["shape_misses"] - st0["shape_misses"] == 1 and st["entries"] >= 1 and st["bytes"] > 0
    ctx.cache_configure(False, 8 << 20)       # Cache.usecache := false: nothing is served
    assert ctx.cache_getshape(41) is None
    ctx.cache_configure(True, 50 << 20)
    ctx.cache_clear()
    assert ctx.cache_stats()["entries"] == 0


def test_partial_sprite_cache(ctx, oracle):
    """The partial-sprite cache (render.ml:1169-1242; cache.ml:328-367, 390-407): a top-level Group with an id keeps a
    sprite that GROWS as more of it is exposed — three different update shapes, each rendering only what is not cached
    yet — and then follows the group through a drag as an alias (the cached sprite is read translated).  Against the
    oracle with Cache.usecache = true and a persistent cache; and the same frames with the cache off."""
    W, H = 400, 300

    def tri(b, x, y, c, a):
        return b.polygon([(x, y), (x + 130.5, y + 24.2), (x + 50.1, y + 140.7)], S.Fill.plain(S.dissolve(c, a)))

    b = S.SceneBuilder()
    b.polygon([(150.2, 20.1), (260.5, 60.9), (170.0, 130.3)], S.Fill.plain(S.dissolve(S.rgba8(250, 220, 30), 150)), oid=5)   # in front of the group
    g = b.group_begin(oid=8)
    tri(b, 40.3, 35.1, S.rgba8(220, 40, 40), 255)
    tri(b, 90.9, 60.4, S.rgba8(40, 220, 40), 170)
    b.rectangle(S.rgba8(30, 30, 160), 120.0, 100.0, 200.0, 170.0)
    tri(b, 60.2, 82.8, S.rgba8(40, 40, 220), 90)
    b.group_end()
    for i in range(6):   # behind it
        x, y = 20.0 + 55 * i, 30.0 + 31 * i
        b.polygon([(x, y), (x + 90.5, y + 10.2), (x + 70.1, y + 95.5), (x - 5.0, y + 60.0)], S.Fill.plain(S.dissolve(S.rgba8(20 * i % 255, 200, 255 - 9 * i), 255 if i % 2 else 170)), oid=100 + i)
    b.begin_background()
    b.rectangle(S.LIGHTGREY, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    group_index = 1
    updates = [util.flat_of_rows([(y, [(30, 90)]) for y in range(40, 120)]),
               util.flat_of_rows([(y, [(60, 50), (150, 120)]) for y in range(90, 200)]),
               util.flat_of_rows([(y, [(0, W)]) for y in range(H)])]
    for usecache in (True, False):
        ctx.cache_clear()
        ctx.cache_configure(usecache, 64 << 20)
        ctx.fb_configure(W, H)
        objs[group_index].dx = objs[group_index].dy = 0
        sc = ctx.scene_create(objs, nbg, edges, points)
        with oracle.Renderer(usecache=usecache) as ref_r:
            try:
                ref_fb = np.zeros((H, W), dtype=np.uint32)
                for k, flat in enumerate(updates):   # the cached sprite grows: 1st partial, 2nd overlapping the 1st, 3rd everything
                    hu = ctx.shape_import(flat)
                    ctx.render_frame_shape(sc, hu)
                    got = ctx.fb_read_sprite(hu)
                    ctx.shape_free(hu)
                    img = ref_r.frame(objs, n - nbg, nbg, edges, points, flat, (0, 0, W, H))
                    want = np.concatenate([img[yy, xx : xx + l] for yy, spans in util.rows_of_flat(flat) for xx, l in spans])
                    assert np.array_equal(got, want), f"update {k} (usecache={usecache})"
                    for yy, spans in util.rows_of_flat(flat):
                        for xx, l in spans:
                            ref_fb[yy, xx : xx + l] = img[yy, xx : xx + l]
                st = ctx.cache_sprite_stats(sc)
                if usecache:
                    assert st["entries"] == 1 and st["bytes"] > 0 and st["sprite_fills"] >= 1
                else:
                    assert st["sprite_fills"] == 0 and st["sprite_hits"] == 0
                tx = ty = 0
                for step, (dx, dy) in enumerate([(5, 3), (-9, 7), (20, -6), (3, 3), (-30, 15)]):   # the group follows the pointer as an alias
                    ctx.scene_drag_object(sc, group_index, dx, dy)
                    tx, ty = tx + dx, ty + dy
                    objs[group_index].dx, objs[group_index].dy = tx, ty
                    full = util.flat_of_rows([(y, [(0, W)]) for y in range(H)])
                    ref = ref_r.frame(objs, n - nbg, nbg, edges, points, full, (0, 0, W, H))
                    got = ctx.fb_read_rgba(0, 0, W, H)
                    assert _max_lsb(got, ref) == 0, f"drag step {step} (usecache={usecache})"
                if usecache:
                    st2 = ctx.cache_sprite_stats(sc)
                    assert st2["sprite_hits"] >= 1, "once the sprite is complete, drag frames are served from it"
            finally:
                ctx.scene_free(sc)
    objs[group_index].dx = objs[group_index].dy = 0
    ctx.cache_configure(True, 50 << 20)
    ctx.cache_clear()


def _random_premultiplied(rng, n):
    a = np.array([rng.choice([0, 1, 127, 128, 254, 255, rng.randint(0, 255)]) for _ in range(n)], dtype=np.uint32)
    ch = [np.array([rng.randint(0, int(v)) for v in a], dtype=np.uint32) for _ in range(3)]
    return ch[0] | (ch[1] << 8) | (ch[2] << 16) | (a << 24)


def test_sprite_operations(ctx, oracle):
    """Sprite.portion / fillshape / sprite_map / map_coords / shape_intersects across the ABI (sprite.mli:96-125), a
    sprite being (shape, RGBA8 per pixel in span order), against the oracle's restatement of sprite.ml."""
    rng = random.Random(21)
    fills = [S.Fill.plain(S.dissolve(S.rgba8(200, 40, 90), 170)),
             S.Fill.gradient((10.0, 20.0), (250.0, 150.0), True, False, S.rgba8(255, 0, 0), S.dissolve(S.rgba8(0, 0, 255), 128)),
             S.Fill.radial((120.0, 80.0), (130.0, 80.0), (220.0, 80.0), False, True, S.rgba8(255, 255, 0), S.rgba8(0, 90, 30))]
    for it in range(6):
        a = util.random_shape_flat(rng)
        b = util.random_shape_flat(rng, x0=20, y0=-10)
        sub = oracle.shape_op("intersection", a, b)
        ha, hb, hsub = ctx.shape_import(a), ctx.shape_import(b), ctx.shape_import(sub)
        assert ctx.shape_intersects(ha, hb) == (len(sub) > 0)
        px = _random_premultiplied(rng, oracle.shape_card(a))
        assert np.array_equal(ctx.sprite_portion(ha, px, hsub), oracle.sprite_portion(a, px, sub))
        if len(oracle.shape_op("difference", b, a)):
            with pytest.raises(abi.CohError):            # Sprite.portion fails unless the shape lies inside the sprite's
                ctx.sprite_portion(ha, px, hb)
        o = abi.CohObject()
        fills[it % 3].apply(o)
        assert np.array_equal(ctx.sprite_fillshape(ha, o), oracle.sprite_fillshape(o, a)), f"fillshape (fill {it % 3})"
        assert np.array_equal(ctx.sprite_map_coords_fill(ha, o, px), oracle.sprite_map_coords_fill(o, a, px)), f"map_coords (fill {it % 3})"
        for op, arg in (("monochrome", 0), ("dissolve", rng.randint(0, 255)), ("dissolve", 0), ("dissolve", 255), ("red_channel", 0), ("green_channel", 0), ("blue_channel", 0)):
            assert np.array_equal(ctx.sprite_map(op, px, arg), oracle.sprite_map(op, px, arg)), op
        for h in (ha, hb, hsub):
            ctx.shape_free(h)
    assert not ctx.shape_intersects(0, 0) and len(ctx.sprite_portion(0, np.zeros(0, np.uint32), 0)) == 0


@pytest.mark.gpu
def test_large_point_arrays_upload_beside_the_object_walk(ctx, oracle):
    """Scenes with 2^20 brush points or more send the points up from a helper thread while coh_scene_create walks the
    objects (host_scene.inl): the frame is the oracle's, and a call that fails gives the early allocation back."""
    W, H = 256, 192
    b = S.SceneBuilder()
    b.brush(0.8, 6.0, [[("C", (20.0, 30.0), (90.0, 170.0), (160.0, 10.0), (230.0, 150.0))]], S.Fill.plain(S.rgba8(200, 40, 40)))
    b.polygon([(30.5, 20.2), (220.1, 60.7), (120.9, 180.3)], S.Fill.plain(S.dissolve(S.rgba8(30, 60, 220), 120)))
    b.begin_background()
    b.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, edges, points = b.arrays()
    pad = np.zeros(((1 << 20) + 7, 2), dtype=np.int32)               # points no object owns, behind the stroke's
    big = np.ascontiguousarray(np.concatenate([np.asarray(points, dtype=np.int32).reshape(-1, 2), pad]))
    ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, big)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    got = np.asarray(ctx.fb_read_rgba(0, 0, W, H)).view(np.uint32).reshape(H, W)
    assert np.array_equal(got, ref)
    ctx.scene_free(sc)
    ctx.sync()
    m0 = ctx.mem_in_use()                                            # (the context's own frame buffers are in place now)
    objs[1].winding = 7                                              # the polygon: no such winding rule
    with pytest.raises(abi.CohError):
        ctx.scene_create(objs, nbg, edges, big)
    ctx.sync()
    assert ctx.mem_in_use() <= m0 + (1 << 20)                        # the 8 MB of points did not stay behind
