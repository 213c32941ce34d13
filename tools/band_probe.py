"""One band of an N-way split of the C2 frame on one GPU (what a rank of the multi-GPU run does), and the C4 drag step,
for different thresholds between the fused walker and the three-phase frame.  usage: python tools/band_probe.py"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coherence_renderer_b200 import abi, bands, scene

def timed(ctx, fn, n=200, warm=10):
    for _ in range(warm): fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(n): fn()
    ctx.sync(); return (time.perf_counter() - t) / n * 1e3

ctx = abi.Context(0)
W, H = 3840, 2160
objs, n, nbg, e, p = scene.lion_scene(W, H, 7.0).arrays()
for thr in (1 << 30, 65536, 4096, 0):
    ctx.set_option("pre_min_pairs", thr)
    row = {"pre_min_pairs": thr}
    for N in (2, 4, 8):
        k = N // 2   # a middle band (the busiest)
        y0, y1 = bands.band_rows(H, N, k)
        ctx.fb_configure(W, H, y0, y1)
        fake = ctx.fb_device_ptr()
        ctx.fb_set_peers([fake])   # mirrored stores (to the same buffer) as in a multi-GPU run: k_prefill off
        sc = ctx.scene_create(objs, nbg, e, p)
        row[f"band_1_of_{N}_ms"] = round(timed(ctx, lambda: ctx.render_frame(sc, (0, 0, W, H))), 4)
        ctx.fb_set_peers([])
        ctx.scene_free(sc)
    ctx.fb_configure(W, H)
    b, mover = scene.drag_scene(W, H, 7.0 * 0.45)
    o2, n2, nbg2, e2, p2 = b.arrays()
    ctx.cache_clear(); ctx.cache_configure(True, 100 << 20)
    sc = ctx.scene_create(o2, nbg2, e2, p2)
    ctx.render_frame(sc, (0, 0, W, H)); ctx.sync()
    f = [0]
    def step():
        f[0] += 1
        ctx.scene_drag_object(sc, mover, round(3 * math.cos(2 * math.pi * f[0] / 250)), round(2 * math.sin(2 * math.pi * f[0] / 250)))
    row["c4_drag_ms"] = round(timed(ctx, step, 500, 20), 4)
    ctx.scene_free(sc)
    print(json.dumps(row), flush=True)
