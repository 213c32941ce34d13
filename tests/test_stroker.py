"""N2 — the stroker (Shapes.strokepath, shapes.ml:203-530).

CPU: the product's host stroker (host_stroke.cpp: std::list rails joined in place in Pdfutil.pair_reduce's order) against the
oracle's restatement (oracle/shapes.hpp: the reference's list code with vectors), bit for bit, on random paths with every
cap and join; and properties that do not depend on either: areas of stroked lines, containment of the centre line, the order
bevel <= round <= mitred of the joins' areas, the circle of a degenerate path.
GPU: coh_strokepath (outline flattened by k_flatten, sorted) against the oracle's edge list, coh_shapeminshape_of_stroke
against the oracle's scan conversion of those edges, and a stroked path rendered as a StrokedPath object in a scene.
"""
import math

import numpy as np
import pytest

from coherence_renderer_b200 import abi


def _random_path(rng, degenerate=True):
    def rp():
        return (float(rng.uniform(0, 300)), float(rng.uniform(0, 220)))

    subpaths = []
    for _ in range(int(rng.integers(1, 4))):
        p, sp = rp(), []
        for _ in range(int(rng.integers(1, 10))):
            if rng.random() < 0.5:
                q = p if (degenerate and rng.random() < 0.08) else rp()   # zero-length lines are cleaned away
                sp.append(("L", p, q))
            else:
                a, b, q = rp(), rp(), rp()
                if degenerate and rng.random() < 0.05:
                    a = p                                                  # ... and so are curves with a doubled end point
                sp.append(("C", p, a, b, q))
            p = q
        subpaths.append(sp)
    return subpaths


def _area(oracle, edges, winding):
    shape, _ = oracle.shapeminshape(edges, winding)
    return oracle.shape_card(shape), shape


def test_host_stroker_equals_oracle_bit_for_bit(oracle):
    rng = np.random.default_rng(2024)
    compared = 0
    for trial in range(150):
        subpaths = _random_path(rng)
        rec, cnt = abi._path_records(subpaths)
        for join in (abi.JOIN_ROUND, abi.JOIN_MITRED, abi.JOIN_BEVEL):
            sc, ec = int(rng.integers(0, 3)), int(rng.integers(0, 3))
            ml, lw = float(rng.choice([1.0, 1.5, 4.0, 10.0])), float(rng.choice([0.5, 2.0, 7.25, 30.0]))
            try:
                ref_out, ref_cnt, ref_w, _ = oracle.strokepath((sc, join, ec, ml, lw), rec, cnt)
            except Exception:
                ref_out = None                      # the reference fails (nothing left of the path after cleaning: pair_reduce [])
            try:
                out, c, w = abi.host_strokepath(abi.strokespec(sc, join, ec, ml, lw), subpaths)
            except abi.CohError:
                out = None
            assert (out is None) == (ref_out is None), (trial, join)
            if out is None:
                continue
            assert out.shape == ref_out.shape and np.array_equal(out.view(np.uint64), ref_out.view(np.uint64)), (trial, join)
            assert np.array_equal(c, ref_cnt) and w == ref_w == abi.COH_EVENODD
            # the outline of every subpath is closed: each segment starts where the one before ended, the last ends at the start
            at = 0
            for n in c:
                o = out[at:at + n]
                end = np.where(o[:, :1] != 0, o[:, 7:9], o[:, 3:5])
                assert np.array_equal(o[1:, 1:3], end[:-1]) and np.array_equal(o[0, 1:3], end[-1]), (trial, join)
                at += n
            compared += 1
    assert compared > 300


def test_stroked_line_areas_and_caps(oracle):
    """A line of length L stroked with width w: L w pixels with butt caps, (L + w) w with projecting caps,
    L w + pi (w / 2)^2 with round caps — between the cardinals of minshape and shape; the centre line lies inside."""
    for (a, b) in [((20.0, 30.0), (220.0, 30.0)), ((30.0, 20.0), (180.0, 170.0)), ((200.5, 40.25), (60.0, 190.0))]:
        L = math.dist(a, b)
        for w in (6.0, 17.5):
            want = {abi.CAP_BUTT: L * w, abi.CAP_PROJECTING: (L + w) * w, abi.CAP_ROUND: L * w + math.pi * (w / 2) ** 2}
            for cap, area in want.items():
                rec, cnt = abi._path_records([[("L", a, b)]])
                _, _, wd, edges = oracle.strokepath((cap, abi.JOIN_BEVEL, cap, 10.0, w), rec, cnt)
                got, shape = _area(oracle, edges, wd)
                inner = oracle.shape_card(oracle.shapeminshape(edges, wd)[1])
                # the shape holds every pixel the antialiased outline can touch, the minshape only pixels well inside it
                assert inner <= area <= got and got - inner <= 6.0 * (L + w) + 40, (a, b, w, cap, inner, area, got)
                rows = {int(r[0]): r[1:] for r in _rows(shape)}
                for t in np.linspace(0.02, 0.98, 25):
                    x, y = int(a[0] + t * (b[0] - a[0])), int(a[1] + t * (b[1] - a[1]))
                    assert y in rows and any(s <= x < s + l for s, l in rows[y].reshape(-1, 2)), (a, b, w, cap, x, y)


def _rows(flat):
    """Rows of a canonical flat span set (y, n, then n (start, length) pairs per row): yields arrays [y, s0, l0, s1, l1, ...]."""
    flat, i = np.asarray(flat), 0
    while i < len(flat):
        k = int(flat[i + 1])
        yield np.concatenate([[flat[i]], flat[i + 2: i + 2 + 2 * k]])
        i += 2 + 2 * k


def test_join_areas_are_ordered_and_mitre_limit_bevels(oracle):
    corner = [[("L", (40.0, 160.0), (140.0, 60.0)), ("L", (140.0, 60.0), (240.0, 160.0))]]   # a right angle
    rec, cnt = abi._path_records(corner)
    area = {}
    for join in (abi.JOIN_BEVEL, abi.JOIN_ROUND, abi.JOIN_MITRED):
        _, _, wd, edges = oracle.strokepath((abi.CAP_BUTT, join, abi.CAP_BUTT, 10.0, 24.0), rec, cnt)
        area[join], _ = _area(oracle, edges, wd)
    w = 24.0
    assert area[abi.JOIN_BEVEL] < area[abi.JOIN_ROUND] < area[abi.JOIN_MITRED]
    # at a right angle the centre, the two rail ends and the mitre's tip make a square of side w / 2: the bevel holds one half
    # of it, the mitre adds the other half, the round join a quarter disc minus the bevel's half
    assert abs((area[abi.JOIN_MITRED] - area[abi.JOIN_BEVEL]) - 0.5 * (w / 2) ** 2) <= 30
    assert abs((area[abi.JOIN_ROUND] - area[abi.JOIN_BEVEL]) - (math.pi / 4 - 0.5) * (w / 2) ** 2) <= 30
    # a mitre limit below 1 / sin (angle / 2) = sqrt 2 turns the mitre into a bevel (shapes.ml:342-346)
    _, _, wd, edges = oracle.strokepath((abi.CAP_BUTT, abi.JOIN_MITRED, abi.CAP_BUTT, 1.2, 24.0), rec, cnt)
    assert _area(oracle, edges, wd)[0] == area[abi.JOIN_BEVEL]


def test_bounds_stroke(oracle):
    """Shapes.bounds_stroke (shapes.ml:522-540): product = oracle, and the box holds the stroke's shape (to within the few pixels of the reference's estimate)."""
    rng = np.random.default_rng(31)
    for trial in range(60):
        subpaths = _random_path(rng, degenerate=False)
        rec, cnt = abi._path_records(subpaths)
        sc, jn, ec = int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(0, 3))
        ml, lw = float(rng.choice([1.0, 2.5, 10.0])), float(rng.choice([0.5, 3.0, 11.5]))
        got = abi.host_bounds_stroke(abi.strokespec(sc, jn, ec, ml, lw), subpaths)
        assert got == oracle.bounds_stroke((sc, jn, ec, ml, lw), rec, cnt), trial
        _, _, wd, edges = oracle.strokepath((sc, jn, ec, ml, lw), rec, cnt)
        shape, _ = oracle.shapeminshape(edges, wd)
        rows = list(_rows(shape))
        if jn == abi.JOIN_MITRED:
            continue   # (a sharp mitre under a small limit is bevelled, under a large one it may still reach beyond: shapes.ml:520-521)
        # the box is the reference's estimate, not a proof: the soft matte and the flaring of offset curves at tight bends
        # (shapes.ml:155-158) reach a few pixels beyond the stroke's half width
        slack = 6
        assert min(int(r[0]) for r in rows) >= got[2] - slack and max(int(r[0]) for r in rows) <= got[3] + slack, trial
        assert min(int(r[1]) for r in rows) >= got[0] - slack and max(int(r[-2] + r[-1] - 1) for r in rows) <= got[1] + slack, trial
    with pytest.raises(abi.CohError):
        abi.host_bounds_stroke(abi.strokespec(0, 0, 0, 10.0, 1.0), [])


def test_degenerate_path_with_round_caps_is_a_circle(oracle):
    rec, cnt = abi._path_records([[("L", (50.0, 60.0), (50.0, 60.0))]])
    out, c, wd, edges = oracle.strokepath((abi.CAP_ROUND, abi.JOIN_BEVEL, abi.CAP_ROUND, 10.0, 30.0), rec, cnt)
    assert wd == abi.COH_NONZERO and list(c) == [4] and np.all(out[:, 0] == 1)
    got, _ = _area(oracle, edges, wd)
    inner = oracle.shape_card(oracle.shapeminshape(edges, wd)[1])
    assert inner <= math.pi * 15 ** 2 <= got and got - inner <= 300   # (minshape inside the disc, shape around it)
    pout, pc, pw = abi.host_strokepath(abi.strokespec(abi.CAP_ROUND, abi.JOIN_BEVEL, abi.CAP_ROUND, 10.0, 30.0), [[("L", (50.0, 60.0), (50.0, 60.0))]])
    assert np.array_equal(pout.view(np.uint64), out.view(np.uint64)) and list(pc) == [4] and pw == wd
    # with any other cap a degenerate path is cleaned away: nothing is stroked (the reference's pair_reduce never runs)
    pout, pc, pw = abi.host_strokepath(abi.strokespec(abi.CAP_BUTT, abi.JOIN_BEVEL, abi.CAP_ROUND, 10.0, 30.0), [[("L", (50.0, 60.0), (50.0, 60.0))]])
    assert len(pout) == 0 and len(pc) == 0


@pytest.mark.gpu
def test_strokepath_on_device(ctx, oracle):
    rng = np.random.default_rng(99)
    for trial in range(24):
        subpaths = _random_path(rng, degenerate=trial % 3 == 0)
        rec, cnt = abi._path_records(subpaths)
        sc, jn, ec = int(rng.integers(0, 3)), trial % 3, int(rng.integers(0, 3))
        ml, lw = float(rng.choice([1.5, 4.0, 10.0])), float(rng.choice([1.0, 5.5, 18.0]))
        try:
            _, _, ref_w, ref_edges = oracle.strokepath((sc, jn, ec, ml, lw), rec, cnt)
        except Exception:
            continue
        spec = abi.strokespec(sc, jn, ec, ml, lw)
        edges, w = ctx.strokepath(spec, subpaths)
        assert w == ref_w and np.array_equal(edges, ref_edges), trial
        s, m = ctx.shapeminshape_of_stroke(spec, subpaths)
        rs, rm = oracle.shapeminshape(ref_edges, ref_w)
        assert np.array_equal(ctx.shape_export(s), rs) and np.array_equal(ctx.shape_export(m), rm), trial
        ctx.shape_free(s)
        ctx.shape_free(m)
    with pytest.raises(abi.CohError):
        ctx.strokepath(abi.strokespec(7, 0, 0, 10.0, 2.0), [[("L", (0.0, 0.0), (5.0, 5.0))]])


@pytest.mark.gpu
def test_stroked_path_objects_in_a_scene(ctx, oracle):
    """Basic (fill, StrokedPath (spec, path)) end to end: the product's stroker gives the object's edges (coh_strokepath),
    the frame renders them with the StrokedPath winding quirk (shape NonZero, sprite EvenOdd: render.ml:510, 1018); the
    oracle strokes the same paths with its own restatement and renders its own edges.  Tight bends and a self-crossing
    path make the outline overlap itself, where the two rules differ."""
    from coherence_renderer_b200 import scene as S

    W, H = 320, 240
    strokes = [
        (abi.strokespec(abi.CAP_ROUND, abi.JOIN_ROUND, abi.CAP_ROUND, 10.0, 14.0),
         [[("L", (30.0, 200.0), (150.0, 40.0)), ("L", (150.0, 40.0), (170.0, 200.0)), ("C", (170.0, 200.0), (300.0, 220.0), (310.0, 20.0), (200.0, 60.0))]],
         S.Fill.plain(S.dissolve(S.rgba8(200, 40, 20), 200))),
        (abi.strokespec(abi.CAP_PROJECTING, abi.JOIN_MITRED, abi.CAP_BUTT, 4.0, 9.5),
         [[("L", (20.0, 30.0), (280.0, 120.0)), ("L", (280.0, 120.0), (40.0, 150.0)), ("L", (40.0, 150.0), (260.0, 35.0))],
          [("C", (60.0, 220.0), (100.0, 160.0), (200.0, 230.0), (250.0, 170.0))]],
         S.Fill.plain(S.rgba8(20, 60, 160))),
        (abi.strokespec(abi.CAP_BUTT, abi.JOIN_BEVEL, abi.CAP_ROUND, 10.0, 3.0),
         [[("C", (10.0, 120.0), (100.0, 10.0), (220.0, 230.0), (310.0, 110.0)), ("L", (310.0, 110.0), (160.0, 118.0))]],
         S.Fill.plain(S.dissolve(S.rgba8(10, 10, 10), 230))),
    ]
    b, bo = S.SceneBuilder(), S.SceneBuilder()
    for spec, path, fill in strokes:
        edges, w = ctx.strokepath(spec, path)
        rec, cnt = abi._path_records(path)
        _, _, ref_w, ref_edges = oracle.strokepath((spec.startcap, spec.join, spec.endcap, spec.mitrelimit, spec.linewidth), rec, cnt)
        assert w == ref_w == abi.COH_EVENODD
        b.stroked_path_edges(edges, fill)
        bo.stroked_path_edges(ref_edges, fill)
    for x in (b, bo):
        x.begin_background()
        x.rectangle(S.WHITE, 0.0, 0.0, float(W), float(H))
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    ctx.render_frame(sc, (0, 0, W, H), abi.COH_RENDER_RECORD_U)
    ctx.sync()
    got = ctx.fb_read_rgba(0, 0, W, H)
    u = ctx.render_uncovered()
    got_u = ctx.shape_export(u)
    ctx.shape_free(u)
    ctx.scene_free(sc)
    objs, n, nbg, e, p = bo.arrays()
    ref, ref_u = oracle.render_frame(objs, n - nbg, nbg, e, p, (0, 0, W, H), want_u=True)
    assert np.array_equal(got_u, ref_u)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert len(np.unique(got)) > 50   # strokes, their antialiased borders and the overlaps are all there
