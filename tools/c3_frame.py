"""Two C3 frames (10^5 random polygons / brush strokes, 7680x4320) through the C ABI — for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
ctx = abi.Context(0)
W, H = 7680, 4320
objs, n, nbg, e, p = scene.random_scene(W, H, 100000).arrays()
ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
for _ in range(2):
    ctx.render_frame(sc, (0, 0, W, H))
ctx.sync()
