"""Shared helpers for the tests: random geometry and flat-shape utilities."""
import math
import random

import numpy as np

from coherence_renderer_b200 import scene as S


def rows_of_flat(flat):
    out, i = [], 0
    flat = [int(v) for v in flat]
    while i < len(flat):
        k = flat[i + 1]
        out.append((flat[i], [(flat[i + 2 + 2 * j], flat[i + 3 + 2 * j]) for j in range(k)]))
        i += 2 + 2 * k
    return out


def flat_of_rows(rows):
    f = []
    for y, spans in rows:
        f += [y, len(spans)]
        for x, l in spans:
            f += [x, l]
    return np.array(f, dtype=np.int32)


def bitmap_of_flat(flat, x0, y0, w, h):
    bm = np.zeros((h, w), dtype=bool)
    for y, spans in rows_of_flat(flat):
        for x, l in spans:
            if y0 <= y < y0 + h:
                a, b = max(x, x0), min(x + l, x0 + w)
                if a < b:
                    bm[y - y0, a - x0 : b - x0] = True
    return bm


def flat_of_bitmap(bm, x0, y0):
    rows = []
    for r in range(bm.shape[0]):
        row = bm[r]
        if not row.any():
            continue
        d = np.diff(np.concatenate(([0], row.astype(np.int8), [0])))
        starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
        rows.append((y0 + r, [(x0 + int(s), int(e - s)) for s, e in zip(starts, ends)]))
    return flat_of_rows(rows)


def random_polygon_edges(rng, lo=-40.0, hi=400.0, rmax=150.0, kmax=10):
    k = rng.randint(3, kmax)
    cx, cy = rng.uniform(lo, hi), rng.uniform(lo, hi)
    R = math.exp(rng.uniform(math.log(2.0), math.log(rmax)))
    if rng.random() < 0.6:
        ang = sorted(rng.uniform(0, 2 * math.pi) for _ in range(k))
    else:
        ang = [rng.uniform(0, 2 * math.pi) for _ in range(k)]  # self-intersecting
    pts = [(cx + R * rng.uniform(0.3, 1) * math.cos(a), cy + R * rng.uniform(0.3, 1) * math.sin(a)) for a in ang]
    if rng.random() < 0.2:
        pts = [(float(round(x)), float(round(y))) for x, y in pts]  # axis-aligned pieces and ties
    s = S.sub_of_float
    edges = [[s(pts[i][0]), s(pts[i][1]), s(pts[(i + 1) % k][0]), s(pts[(i + 1) % k][1])] for i in range(k)]
    if rng.random() < 0.15:
        edges = edges[:-1]  # an unclosed subpath stays open (polygon.ml:222-228)
    return np.array(edges, dtype=np.int32)


def random_shape_flat(rng, x0=-30, y0=-30, w=300, h=200, density=0.3):
    """A random canonical span set inside the given box."""
    rows = []
    for y in range(y0, y0 + h):
        if rng.random() > 0.8:
            continue
        spans, x = [], x0 + rng.randint(0, 20)
        while x < x0 + w:
            l = rng.randint(1, 40)
            if rng.random() < density:
                spans.append((x, min(l, x0 + w - x)))
            x += l + rng.randint(1, 30)
        if spans:
            rows.append((y, spans))
    return flat_of_rows(rows)
