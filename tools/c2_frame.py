"""Three C2 frames (lion, 3840x2160, scale 7) through the C ABI — the command the ncu captures under profiles/ run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
ctx = abi.Context(0)
W, H = 3840, 2160
objs, n, nbg, e, p = scene.lion_scene(W, H, 7.0).arrays()
ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
for _ in range(3):
    ctx.render_frame(sc, (0, 0, W, H))
ctx.sync()
