"""Small end-to-end run for compute-sanitizer (memcheck): lion, random scene with brushes, fancy fills,
Convolved object, shape algebra, update shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coherence_renderer_b200 import abi, scene as S
ctx = abi.Context(0)
def run(b, W, H):
    objs, n, nbg, e, p = b.arrays()
    ctx.fb_configure(W, H); sc = ctx.scene_create(objs, nbg, e, p)
    ctx.render_frame(sc, (0, 0, W, H), abi.COH_RENDER_RECORD_U); ctx.sync()
    u = ctx.render_uncovered(); ctx.shape_export(u); ctx.shape_free(u)
    img = ctx.fb_read_rgba(0, 0, W, H); ctx.scene_free(sc); return img
run(S.lion_scene(333, 257, 0.7), 333, 257)
for opts in ({"fused": 0}, {"fused": 0, "comp_rows": 0}, {"fused": 0, "aa_general": 1}, {"fused": 1, "walk_h": 4}):   # three-phase frame (interval AA + row compositor), its other variants, the fused walker
    for k, v in opts.items(): ctx.set_option(k, v)
    run(S.lion_scene(333, 257, 0.7), 333, 257)
    for k in opts: ctx.set_option(k, {"fused": -1, "comp_rows": 1, "aa_general": 0, "walk_h": 0}[k])
# partial-sprite cache: a group with an id, a partial update, a drag
b = S.SceneBuilder(); S.add_lion(b, 333, 257, 0.7, oid=4); b.begin_background(); b.rectangle(S.LIGHTGREY, 0.0, 0.0, 333.0, 257.0)
objs, n, nbg, e, p = b.arrays()
ctx.fb_configure(333, 257); sc = ctx.scene_create(objs, nbg, e, p)
ctx.render_frame(sc, (40, 30, 120, 90)); ctx.render_frame(sc, (0, 0, 333, 257))
for d in ((3, 2), (-5, 4)): ctx.scene_drag_object(sc, 0, *d)
ctx.sync(); ctx.scene_free(sc)
run(S.random_scene(300, 200, 1200, seed=3), 300, 200)
b = S.SceneBuilder()
b.polygon([(30.5, 30.5), (170.2, 40.1), (150.0, 140.0), (40.0, 120.0)], S.Fill.gradient((20.0, 20.0), (150.0, 120.0), True, False, S.rgba8(255, 0, 0), S.rgba8(0, 0, 255)))
b.polygon([(50.0, 40.0), (150.0, 45.5), (140.5, 120.0)], S.Fill.plain(S.rgba8(0, 0, 0)), convolve=("gaussian", 4))
b.group_begin(convolve=("unit", 2), pretrans=200)   # Convolved (kernel, Group members)
b.polygon([(20.0, 90.0), (120.0, 95.5), (110.5, 150.0)], S.Fill.plain(S.rgba8(0, 90, 0)))
b.rectangle(S.rgba8(200, 30, 30), 60.0, 100.0, 90.0, 130.0)
b.group_end()
b.begin_background(); b.rectangle(S.WHITE, 0.0, 0.0, 200.0, 160.0)
run(b, 200, 160)
a = ctx.shape_box(5, 5, 100, 50); c = ctx.shape_box(50, 20, 100, 70)
for f in (ctx.shape_union, ctx.shape_difference, ctx.shape_intersection): ctx.shape_free(f(a, c))
ctx.shape_free(ctx.shape_bloat(a, 3, 2)); ctx.shape_free(ctx.shape_erode(a, 3, 2))
spec = abi.strokespec(abi.CAP_ROUND, abi.JOIN_ROUND, abi.CAP_PROJECTING, 10.0, 7.5)   # the stroker: outline -> k_flatten -> scan conversion
path = [[("L", (20.0, 20.0), (120.0, 30.0)), ("C", (120.0, 30.0), (160.0, 90.0), (60.0, 130.0), (30.0, 80.0))]]
ctx.strokepath(spec, path)
for h in ctx.shapeminshape_of_stroke(spec, path): ctx.shape_free(h)
print("sanitize run OK")
