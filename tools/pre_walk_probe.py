"""C5 and C4 frame times for different thresholds of the three-phase path on passes that composite with the walker."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene

def timed(ctx, fn, n=100, warm=5):
    for _ in range(warm): fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(n): fn()
    ctx.sync(); return (time.perf_counter() - t) / n * 1e3

ctx = abi.Context(0)
W, H = 3840, 2160
o5 = scene.filter_scene(W, H, 7.0).arrays()
b, mover = scene.drag_scene(W, H, 7.0 * 0.45)
o4 = b.arrays()
for thr in (1 << 30, 262144, 65536, 16384, 4096, 0):
    ctx.set_option("pre_min_pairs_walk", thr)
    row = {"pre_min_pairs_walk": thr}
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(o5[0], o5[2], o5[3], o5[4])
    l0 = ctx.launch_count(); ctx.render_frame(sc, (0, 0, W, H)); ctx.sync(); row["c5_launches"] = ctx.launch_count() - l0
    row["c5_ms"] = round(timed(ctx, lambda: ctx.render_frame(sc, (0, 0, W, H)), 30), 4)
    ctx.scene_free(sc)
    for cache in (True, False):
        ctx.cache_clear(); ctx.cache_configure(cache, 100 << 20)
        sc = ctx.scene_create(o4[0], o4[2], o4[3], o4[4])
        ctx.render_frame(sc, (0, 0, W, H)); ctx.sync()
        f = [0]
        def step():
            f[0] += 1
            ctx.scene_drag_object(sc, mover, round(3 * math.cos(2 * math.pi * f[0] / 250)), round(2 * math.sin(2 * math.pi * f[0] / 250)))
        row["c4_drag_ms" + ("" if cache else "_nocache")] = round(timed(ctx, step, 300, 20), 4)
        ctx.scene_free(sc)
    print(json.dumps(row), flush=True)
