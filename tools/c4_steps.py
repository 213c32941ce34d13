"""Dev tool: a few fused C4 drag steps (coh_scene_drag_object) — run under ncu for the launch list of one step."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
b, mover = scene.drag_scene(W, H, 7.0 * 0.45)
objs, n, nbg, e, p = b.arrays()
ctx = abi.Context(0)
ctx.cache_clear(); ctx.cache_configure(True, 100 << 20); ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
ctx.render_frame(sc, (0, 0, W, H)); ctx.sync()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for f in range(steps):
    if f == steps - 2:
        ctx.sync(); l0 = ctx.launch_count(); t = time.perf_counter()
    ctx.scene_drag_object(sc, mover, round(3 * math.cos(2 * math.pi * f / 250)) or 1, round(2 * math.sin(2 * math.pi * f / 250)) or 1)
ctx.sync()
print("last 2 steps: %.4f ms/step, %d launches/step" % ((time.perf_counter() - t) * 1e3 / 2, (ctx.launch_count() - l0) // 2))
