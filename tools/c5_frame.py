"""Dev tool: render the C5 (filter-heavy) frame a few times — for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
objs, n, nbg, e, p = scene.filter_scene(W, H, 7.0).arrays()
ctx = abi.Context(0); ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.render_frame(sc, (0, 0, W, H))
ctx.sync()
print("ok")
