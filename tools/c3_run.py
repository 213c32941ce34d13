"""Dev tool: C3 (10^5 random polygons / brush strokes at 7680x4320) timing on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
W, H, N = 7680, 4320, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t = time.time(); objs, n, nbg, edges, points = scene.random_scene(W, H, N).arrays(); print("scene build s", round(time.time() - t, 1), "edges", len(edges), "points", len(points))
ctx = abi.Context(0); ctx.fb_configure(W, H)
t = time.time(); sc = ctx.scene_create(objs, nbg, edges, points); print("scene_create s", round(time.time() - t, 2))
for i in range(2): ctx.render_frame(sc, (0, 0, W, H)); ctx.sync()
ctx.set_timing(True)
t = time.time()
for i in range(5): ctx.render_frame(sc, (0, 0, W, H))
ctx.sync(); dt = (time.time() - t) / 5
print("frame ms", dt * 1e3, "timing (walk, bin, n)", ctx.get_timing())
img = ctx.fb_read_rgba(0, 0, W, H); print("checksum", int(img[::7, ::5].astype("uint64").sum()), "alpha255", float((img >> 24 == 255).mean()))
