"""Differential test of the two restatements of the reference: tests/independent_ref.py (plain Python lists, written
from polygon.ml / colour.ml / coord.ml directly) against oracle/ (C++, the checker of the GPU parity tests).  They share
no code; a transcription error in either shows up as a disagreement.  Also: the hand-derived known answers of
SURVEY.md Appendix C through the independent restatement."""
import json
import os
import random

import numpy as np

from tests import independent_ref as R
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))


def _small_edges(rng):
    """Edge lists that stay cheap for the pure-Python restatement: few rows, all the awkward cases (self-intersecting,
    axis-aligned pieces and ties, unclosed subpaths, negative coordinates)."""
    return util.random_polygon_edges(rng, lo=-12.0, hi=44.0, rmax=18.0, kmax=9)


def test_shape_and_minshape_on_ten_thousand_edge_lists(oracle):
    rng = random.Random(20261018)
    nonempty = 0
    for it in range(10000):
        edges = _small_edges(rng)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        s, m = R.shapeminshape_rows([tuple(int(v) for v in e) for e in edges], w)
        assert R.flat_of_rows(s) == [int(v) for v in ref_s], f"shape differs (case {it}, winding {w}): {edges.tolist()}"
        assert R.flat_of_rows(m) == [int(v) for v in ref_m], f"minshape differs (case {it}, winding {w}): {edges.tolist()}"
        nonempty += bool(len(ref_m))
    assert nonempty > 3000


def test_antialiased_opacity(oracle):
    rng = random.Random(7)
    checked = 0
    for it in range(120):
        edges = _small_edges(rng)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        if len(ref_s) == 0:
            continue
        el = [tuple(int(v) for v in e) for e in edges]
        assert R.flat_of_rows(R.scaled_shape_rows(el, w)) == [int(v) for v in oracle.scaled_shape(edges, w)], f"x16 shape differs (case {it})"
        scaled = R.scaled_shape_rows(el, w)
        rows = util.rows_of_flat(ref_s)
        for y, spans in rng.sample(rows, min(2, len(rows))):
            x0, l = spans[0]
            n = min(l, 12)
            ref = oracle.polygon_opacity(edges, w, [y, 1, x0, n])
            got = [R.pixel_opacity(scaled, x0 + k, y) for k in range(n)]
            assert got == [int(v) for v in ref], f"opacity differs (case {it}, row {y})"
            checked += n
    assert checked > 800


def test_colour_operators(oracle):
    rng = random.Random(34)

    def rnd():
        a = rng.choice([0, 1, 127, 128, 254, 255, rng.randint(0, 255)])
        return (rng.randint(0, a), rng.randint(0, a), rng.randint(0, a), a)

    def word(c):
        return c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24)

    for i in range(65026):
        assert R.div255(i) == oracle.div255(i) == i // 255
    for _ in range(20000):
        a, b, d = rnd(), rnd(), rng.randint(0, 255)
        assert word(R.over(a, b)) == oracle.colour_op("over", word(a), word(b))
        assert word(R.dissolve(a, d)) == oracle.colour_op("dissolve", word(a), d)
        assert R.alpha_over(a, b)[3] == oracle.colour_op("alpha_over", word(a), word(b)) >> 24
        da, db = R.dissolve(a, d), R.dissolve(b, 255 - d)
        if all(x + y <= 255 for x, y in zip(da, db)):
            assert word(R.dissolve_between(a, b, d)) == oracle.colour_op("dissolve_between", word(a), word(b), d)


def test_appendix_c_known_answers():
    kc = json.load(open(os.path.join(HERE, "golden", "appendix_c.json")))
    assert R.VOLUME == kc["aa"]["volume"] and sum(sum(r) for r in R.MAINTABLE) == kc["aa"]["maintable_sum"]
    assert R.MAINTABLE[15][15] == kc["aa"]["centre"] and R.MAINTABLE[0][0] == kc["aa"]["corner"]
    k = kc["rect_10_10_20_20_evenodd"]
    rect = [tuple(e) for e in k["edges"]]
    s, m = R.shapeminshape_rows(rect, 1)
    assert sorted(s) == list(range(9, 21)) and all(s[y] == set(range(9, 21)) for y in s)
    assert sorted(m) == list(range(11, 19)) and all(m[y] == set(range(11, 19)) for y in m)
    scaled = R.scaled_shape_rows(rect, 1)
    assert sorted(scaled) == list(range(152, 314)) and all(scaled[y] == set(range(152, 314)) for y in scaled)
    grid = k["opacity_x_8_to_22"]
    for y, row in grid.items():
        assert [R.pixel_opacity(scaled, x, int(y)) for x in range(8, 23)] == row, f"opacity row {y}"
