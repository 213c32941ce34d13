"""The oracle against the hand-derived known answers of SURVEY.md Appendix C (tests/golden/appendix_c.json)
and against exhaustive / algebraic properties of the colour codec and operators (colour.ml)."""
import json
import os
import random

import numpy as np

from tests import util

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "appendix_c.json")))


def test_colour_constants(oracle):
    c = G["colours"]
    assert oracle.colour_of_rgba(0, 0, 0, 0) == int(c["clear"], 16)
    assert oracle.colour_of_rgba(255, 255, 255, 255) == int(c["white"], 16)
    assert oracle.colour_of_rgba(0, 0, 0, 255) == int(c["black"], 16)
    assert oracle.colour_of_rgba(255, 0, 0, 255) == int(c["red"], 16)
    assert oracle.colour_of_rgba(10, 20, 30, 200) == int(c["rgba_10_20_30_200"], 16)


def test_colour_codec_roundtrip_premultiplied(oracle):
    rng = random.Random(5)
    for a in list(range(0, 256, 5)) + [1, 2, 254, 255]:
        for _ in range(60):
            r, g, b = (rng.randint(0, a) for _ in range(3))
            w = r | (g << 8) | (b << 16) | (a << 24)
            col = oracle.colour_of_rgba8(w)
            assert 0 <= col < 2 ** 31
            assert oracle.rgba8_of_colour(col) == w
    # edge rows of the pyramid: channels equal to alpha, and alpha +-1 around the 7-bit split
    for a in range(256):
        for r, g, b in ((a, a, a), (a, 0, 0), (0, a, 0), (0, 0, a), (max(a - 1, 0), a, max(a - 1, 0))):
            w = r | (g << 8) | (b << 16) | (a << 24)
            assert oracle.rgba8_of_colour(oracle.colour_of_rgba8(w)) == w


def test_div255_exact(oracle):
    for i in range(0, 65026):
        assert oracle.div255(i) == i // 255


def test_over_and_dissolve_properties(oracle):
    rng = random.Random(6)

    def rnd():
        a = rng.randint(0, 255)
        return rng.randint(0, a) | (rng.randint(0, a) << 8) | (rng.randint(0, a) << 16) | (a << 24)

    for _ in range(3000):
        a, b = rnd(), rnd()
        assert oracle.colour_op("over", 0, b) == b                     # alpha 0 -> b (colour.ml:316)
        opaque = (a & 0x00FFFFFF) | 0xFF000000
        assert oracle.colour_op("over", opaque, b) == opaque            # alpha 255 -> a
        o = oracle.colour_op("over", a, b)
        ch = [(o >> s) & 255 for s in (0, 8, 16, 24)]
        assert all(c <= ch[3] for c in ch[:3])                          # stays premultiplied
        d = rng.randint(0, 255)
        ds = oracle.colour_op("dissolve", a, d)
        assert [(ds >> s) & 255 for s in (0, 8, 16, 24)] == [((a >> s) & 255) * d // 255 if 0 < d < 255 else (0 if d == 0 else (a >> s) & 255) for s in (0, 8, 16, 24)]
        assert oracle.colour_op("dissolve_between", a, b, 255) == a and oracle.colour_op("dissolve_between", a, b, 0) == b


def test_aa_table(oracle):
    M, vol = oracle.aa_tables()
    assert int(M.sum()) == G["aa"]["maintable_sum"] and vol == G["aa"]["volume"]
    assert int(M[15, 15]) == G["aa"]["centre"] and int(M[0, 0]) == G["aa"]["corner"]
    assert (256 * int(M.sum()) + vol // 2) // vol == G["aa"]["full_window_opacity"]
    assert np.array_equal(M, M.T)


def test_rectangle_known_answer(oracle):
    R = G["rect_10_10_20_20_evenodd"]
    assert oracle.sub_of_float(10.0) == G["sub_of_float_10"]
    shp, mshp = oracle.shapeminshape(R["edges"], 1)
    rows = util.rows_of_flat(shp)
    assert [rows[0][0], rows[-1][0]] == R["shape_rows"] and all(sp == [tuple(R["shape_span"])] for _, sp in rows)
    mrows = util.rows_of_flat(mshp)
    assert [mrows[0][0], mrows[-1][0]] == R["minshape_rows"] and all(sp == [tuple(R["minshape_span"])] for _, sp in mrows)
    srows = util.rows_of_flat(oracle.scaled_shape(R["edges"], 1))
    assert [srows[0][0], srows[-1][0]] == R["scaled_rows"] and all(sp == [tuple(R["scaled_span"])] for _, sp in srows)
    for y, want in R["opacity_x_8_to_22"].items():
        got = oracle.polygon_opacity(R["edges"], 1, [int(y), 1, 8, 15])
        assert got.tolist() == want, f"row {y}"
    # shifting the rectangle by (+0.5, +0.5) shifts both shapes by exactly one pixel
    s = oracle.sub_of_float
    pts = [(10.5, 10.5), (20.5, 10.5), (20.5, 20.5), (10.5, 20.5)]
    e = [[s(pts[i][0]), s(pts[i][1]), s(pts[(i + 1) % 4][0]), s(pts[(i + 1) % 4][1])] for i in range(4)]
    shp, mshp = oracle.shapeminshape(e, 1)
    Q = G["rect_shifted_half"]
    rows, mrows = util.rows_of_flat(shp), util.rows_of_flat(mshp)
    assert [rows[0][0], rows[-1][0]] == Q["shape_rows"] and rows[0][1] == [tuple(Q["shape_span"])]
    assert [mrows[0][0], mrows[-1][0]] == Q["minshape_rows"] and mrows[0][1] == [tuple(Q["minshape_span"])]


def test_triangle_known_answer(oracle):
    T = G["triangle_nonzero"]
    s = oracle.sub_of_float
    p = T["points"]
    e = [[s(p[i][0]), s(p[i][1]), s(p[(i + 1) % 3][0]), s(p[(i + 1) % 3][1])] for i in range(3)]
    shp, mshp = oracle.shapeminshape(e, 0)
    rows, mrows = dict(util.rows_of_flat(shp)), dict(util.rows_of_flat(mshp))
    for y, spans in T["shape_first_rows"].items():
        assert rows[int(y)] == [tuple(x) for x in spans]
    assert max(rows) == T["shape_last_row"]
    for y, spans in {**T["minshape_first_rows"], **T["minshape_last"]}.items():
        assert mrows[int(y)] == [tuple(x) for x in spans]
    assert max(mrows) == 22
