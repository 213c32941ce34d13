"""The product's per-lane row arithmetic (coherence_renderer_b200/csrc/raster_core.cuh — the code the CUDA
kernels execute) compiled for the host by tests/host_emul and checked against the oracle.  This is a test
harness only; the product library has no CPU path.  The GPU runs of the same code are in test_gpu_parity.py."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from coherence_renderer_b200 import scene as S
from tests import util

E = C.CDLL(os.path.join(os.path.dirname(__file__), "host_emul", "libemul.so"))
for f in ("emul_over", "emul_dissolve", "emul_dissolve_between", "emul_alpha_over", "emul_fill"):
    getattr(E, f).restype = C.c_uint32


def emul_shapes(edges, winding, chunk_words):
    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 4)
    ps, pm, ns, nm = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64(), C.c_int64()
    rc = E.emul_shapeminshape(e.ctypes.data_as(C.POINTER(C.c_int32)), len(e), winding, chunk_words, C.byref(ps), C.byref(ns), C.byref(pm), C.byref(nm))
    assert rc == 0, f"emul_shapeminshape rc={rc} (3/4: a pixel fell outside the product's conservative object box)"
    a = np.ctypeslib.as_array(ps, shape=(max(ns.value, 1),))[: ns.value].copy()
    b = np.ctypeslib.as_array(pm, shape=(max(nm.value, 1),))[: nm.value].copy()
    E.emul_free(ps)
    E.emul_free(pm)
    return a, b


@pytest.mark.parametrize("chunk_words", [0, 1, 3])
def test_scan_row_windows_against_oracle(oracle, chunk_words):
    """chunk_words = 1 is the walker's 32-pixel tile window (windowed winding with left / right groups)."""
    rng = random.Random(31 + chunk_words)
    for it in range(700):
        edges = util.random_polygon_edges(rng)
        w = rng.randint(0, 1)
        ref_s, ref_m = oracle.shapeminshape(edges, w)
        got_s, got_m = emul_shapes(edges, w, chunk_words)
        assert np.array_equal(got_s, ref_s) and np.array_equal(got_m, ref_m), f"case {it}"


def test_scan_row_lion_objects(oracle):
    b = S.lion_scene(1280, 1024, 3.0)
    objs, n, nbg, edges, points = b.arrays()
    for i in range(n):
        o = objs[i]
        if o.kind != 0:
            continue
        e = edges[o.first : o.first + o.count]
        ref_s, ref_m = oracle.shapeminshape(e, o.winding)
        got_s, got_m = emul_shapes(e, o.winding, 1)
        assert np.array_equal(got_s, ref_s) and np.array_equal(got_m, ref_m), f"lion object {i}"


@pytest.mark.parametrize("fast", [False, True])
def test_aa_opacity_words_against_oracle(oracle, fast):
    """aa_tile's arithmetic (32 super-sampled rows -> 544-bit rows -> prefix-table sums), general path and the
    register fast path with its fallback."""
    rng = random.Random(33)
    fn = E.emul_opacity_word_fast if fast else E.emul_opacity_word
    words = 0
    for it in range(120):
        edges = util.random_polygon_edges(rng, lo=-20, hi=200, rmax=60, kmax=8)
        w = rng.randint(0, 1)
        shp, _ = oracle.shapeminshape(edges, w)
        rows = util.rows_of_flat(shp)
        if not rows:
            continue
        xmin = min(x for _, sp in rows for x, _ in sp)
        xmax = max(x + l for _, sp in rows for x, l in sp)
        for y, _ in rng.sample(rows, min(3, len(rows))):
            x0 = xmin - 2
            while x0 <= xmax + 2:
                ref = oracle.polygon_opacity(edges, w, [y, 1, x0, 32])
                out = np.zeros(32, dtype=np.uint8)
                rc = fn(edges.ctypes.data_as(C.POINTER(C.c_int32)), len(edges), w, int(x0), int(y), out.ctypes.data_as(C.POINTER(C.c_uint8)))
                assert rc >= 0 and np.array_equal(ref, out), f"case {it} word at x={x0}, y={y}"
                words += 1
                x0 += 32
    assert words > 300


def test_colour_ops_against_oracle(oracle):
    rng = random.Random(34)

    def rnd():
        a = rng.choice([0, 1, 127, 128, 254, 255, rng.randint(0, 255)])
        return rng.randint(0, a) | (rng.randint(0, a) << 8) | (rng.randint(0, a) << 16) | (a << 24)

    for _ in range(5000):
        a, b, d = rnd(), rnd(), rng.randint(0, 255)
        assert E.emul_over(C.c_uint32(a), C.c_uint32(b)) == oracle.colour_op("over", a, b)
        assert E.emul_dissolve(C.c_uint32(a), d) == oracle.colour_op("dissolve", a, d)
        aa, ab = a >> 24, b >> 24
        assert E.emul_alpha_over(C.c_uint32(aa), C.c_uint32(ab)) == oracle.colour_op("alpha_over", a, b) >> 24
        da, db = oracle.colour_op("dissolve", a, d), oracle.colour_op("dissolve", b, 255 - d)
        if all(((da >> s) & 255) + ((db >> s) & 255) <= 255 for s in (0, 8, 16, 24)):
            assert E.emul_dissolve_between(C.c_uint32(a), C.c_uint32(b), d) == oracle.colour_op("dissolve_between", a, b, d)


def test_fills_against_oracle(oracle):
    from coherence_renderer_b200.abi import CohObject

    rng = random.Random(35)
    for _ in range(40):
        kind = rng.choice([1, 2])
        cs, ce = S.dissolve(S.rgba8(rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255)), rng.randint(1, 255)), S.rgba8(rng.randint(0, 255), rng.randint(0, 255), rng.randint(0, 255))
        flags = rng.randint(0, 3)
        p = [rng.uniform(0, 100) for _ in range(6)]
        if rng.random() < 0.15:
            p[2], p[3] = p[0], p[1]  # degenerate axis / zero inner radius
        o = CohObject()
        o.fill_kind, o.colour0, o.colour1, o.fill_flags = kind, cs, ce, flags
        for i in range(6):
            o.fparam[i] = p[i]
        arr = (C.c_double * 6)(*p)
        # the oracle evaluates fills through polygon_sprite on one-pixel spans with full coverage
        for _ in range(60):
            x, y = rng.randint(-20, 130), rng.randint(-20, 130)
            mine = E.emul_fill(kind, C.c_uint32(cs), C.c_uint32(ce), flags, arr, x, y)
            big = [[(x - 40) * 32, (y - 40) * 32, (x + 40) * 32, (y - 40) * 32], [(x + 40) * 32, (y - 40) * 32, (x + 40) * 32, (y + 40) * 32],
                   [(x + 40) * 32, (y + 40) * 32, (x - 40) * 32, (y + 40) * 32], [(x - 40) * 32, (y + 40) * 32, (x - 40) * 32, (y - 40) * 32]]
            ref = oracle.polygon_sprite(o, big, 0, [y, 1, x, 1])
            assert mine == int(ref[0])
