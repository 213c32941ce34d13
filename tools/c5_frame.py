import sys, os
sys.path.insert(0, '/root/repo')
from coherence_renderer_b200 import abi, scene
ctx = abi.Context(0)
W, H = 3840, 2160
objs, n, nbg, e, p = scene.filter_scene(W, H, 7.0).arrays()
ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
for _ in range(3):
    ctx.render_frame(sc, (0, 0, W, H))
ctx.sync()
