"""Per-kernel times of one band of an N-way split of the C2 frame (run under `ncu --metrics gpu__time_duration.sum`):
three whole frames, then three frames of band k of N.  usage: python tools/band_kernels.py N k"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, bands, scene
N, k = int(sys.argv[1]), int(sys.argv[2])
ctx = abi.Context(0)
W, H = 3840, 2160
objs, n, nbg, e, p = scene.lion_scene(W, H, 7.0).arrays()
for (y0, y1) in ((0, H), bands.band_rows(H, N, k)):
    ctx.fb_configure(W, H, y0, y1)
    sc = ctx.scene_create(objs, nbg, e, p)
    for _ in range(3):
        ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    ctx.scene_free(sc)
print("ok")
