"""BASELINE.json's configurations C3, C4 and C5 AT THEIR STATED SIZES, CUDA path against the oracle.

The oracle is a single-threaded CPU restatement, so whole frames of these sizes are out of reach
for it inside a test; it renders horizontal bands (update = Sprite.box 0 y0 W h) of the very same
scene instead and the GPU frame must agree on those rows bit for bit — RGBA and the covered-so-far
set `u` (render.ml:1308).  The oracle's band is rendered with the trivial reject of
render.ml:1270-1279 switched off: that reject tests an object's `bounds` against the bounding box of
the WHOLE current u, a frame-global quantity, so the reference itself renders a band differently
from the same rows of a whole frame in the rare case the reject is not neutral (DESIGN.md §6;
tests/test_bands_gloo.py documents it).
"""
import math

import numpy as np
import pytest

from coherence_renderer_b200 import abi, scene as S

pytestmark = pytest.mark.gpu


def _max_lsb(a, b):
    return int(np.abs(a.view(np.uint8).astype(int) - b.view(np.uint8).astype(int)).max())


def _bands_against_oracle(ctx, oracle, arrays, W, H, bands, check_u=True):
    objs, n, nbg, edges, points = arrays
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    try:
        # the whole frame in one go (the code path the benchmark of this configuration runs) ...
        ctx.render_frame(sc, (0, 0, W, H))
        ctx.sync()
        full = {y0: ctx.fb_read_rgba(0, y0, W, h) for y0, h in bands}
        for y0, h in bands:
            ref, ref_u = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, y0, W, h), bbox_reject=False, want_u=True)
            assert _max_lsb(full[y0], ref) == 0, f"rows {y0}..{y0 + h - 1} of the full frame differ from the oracle"
            # ... and the band on its own as the update (small launch, different walker variant), with `u`
            ctx.render_frame(sc, (0, y0, W, h), abi.COH_RENDER_RECORD_U)
            ctx.sync()
            assert _max_lsb(ctx.fb_read_rgba(0, y0, W, h), ref) == 0, f"band update {y0}..{y0 + h - 1} differs from the oracle"
            if check_u:
                hu = ctx.render_uncovered()
                got_u = ctx.shape_export(hu)
                ctx.shape_free(hu)
                assert np.array_equal(got_u, ref_u), f"covered-so-far set of rows {y0}..{y0 + h - 1} differs"
    finally:
        ctx.scene_free(sc)


def test_c3_full_size_bands(ctx, oracle):
    """C3: 10^5 random layered polygons / brush strokes at 7680x4320 (seed 0xC0FFEE): three 64-row bands."""
    W, H = 7680, 4320
    arrays = S.random_scene(W, H, 100000).arrays()
    _bands_against_oracle(ctx, oracle, arrays, W, H, [(1000, 64), (2160, 64), (4000, 64)])


def test_c5_full_size_bands(ctx, oracle):
    """C5: blur / monochrome / affine lenses over the lion + a Convolved page shadow at 3840x2160: bands through every
    lens (blur + monochrome; blur's lower rim + affine's upper edge; affine) and both horizontal edges of the shadow."""
    W, H = 3840, 2160
    arrays = S.filter_scene(W, H, 7.0).arrays()
    _bands_against_oracle(ctx, oracle, arrays, W, H, [(150, 48), (700, 48), (1236, 48), (1500, 48), (1960, 48)])


def test_c4_full_size_drag(ctx, oracle):
    """C4: 1000-frame drag of the lion group over 400 static objects at 3840x2160, cache on (100 MiB, engine.ml:1610):
    every frame only re-renders the dirty region formed from HBM-resident span sets; every 100th frame the
    incrementally maintained framebuffer must equal the oracle's full render of the moved scene."""
    W, H = 3840, 2160
    b, mover = S.drag_scene(W, H, 7.0 * 0.45)
    objs, n, nbg, edges, points = b.arrays()
    ctx.cache_clear()
    ctx.cache_configure(True, 100 << 20)
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, edges, points)
    try:
        ctx.render_frame(sc, (0, 0, W, H))
        members = []
        k, depth = mover + 1, 1
        while depth:
            if objs[k].kind == abi.COH_OBJ_GROUP_BEGIN:
                depth += 1
            elif objs[k].kind == abi.COH_OBJ_GROUP_END:
                depth -= 1
            else:
                members.append(k)
            k += 1
        tx = ty = 0
        for f in range(1000):
            dx, dy = round(3 * math.cos(2 * math.pi * f / 250)), round(2 * math.sin(2 * math.pi * f / 250))
            ctx.scene_drag_object(sc, mover, dx, dy)
            tx, ty = tx + dx, ty + dy
            if f % 100 == 99 or f == 0:
                ctx.sync()
                got = ctx.fb_read_rgba(0, 0, W, H)
                for m in members:
                    objs[m].dx, objs[m].dy = tx, ty
                ref = oracle.render_frame(objs, n - nbg, nbg, edges, points, (0, 0, W, H))
                assert _max_lsb(got, ref) == 0, f"frame {f} of the drag differs from the oracle's full render"
        st = ctx.cache_sprite_stats(sc)
        assert st["entries"] == 1 and st["bytes"] > 0 and st["sprite_hits"] > 900, "the drag must run from the lion's cached sprite"
    finally:
        ctx.scene_free(sc)
        ctx.cache_clear()
