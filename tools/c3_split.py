"""Dev tool: C3 timing with and without brush strokes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
W, H, N = 7680, 4320, 100000
ctx = abi.Context(0); ctx.fb_configure(W, H)
for bf in (0.0, 0.2, 1.0):
    n_obj = N if bf < 1.0 else 20000
    objs, n, nbg, edges, points = scene.random_scene(W, H, n_obj, brush_fraction=bf).arrays()
    sc = ctx.scene_create(objs, nbg, edges, points)
    for i in range(2): ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync(); ctx.set_timing(True)
    for i in range(5): ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    print("brush_fraction", bf, "objects", n_obj, "edges", len(edges), "points", len(points), "timing (walk, bin, n)", ctx.get_timing())
    ctx.set_timing(False); ctx.scene_free(sc)
