"""Secondary configurations of BASELINE.json on one GPU (the headline C2 number is bench.py's):
  C3  10^5 random layered polygons / brush strokes at 7680x4320 (full frames)
  C4  1000-frame drag of the lion group over 400 static objects at 3840x2160: per frame the group becomes an alias
      of its cached self, the dirty region comes from HBM-resident span sets, only that region is re-rendered
  C5  filter-heavy 4K frame: blur / monochrome / affine lenses over the lion + a Convolved page shadow
Prints one JSON line per configuration.  usage: python tools/bench_configs.py [c3] [c4] [c5] [--frames N]"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from coherence_renderer_b200 import abi, scene  # noqa: E402


def timed_frames(ctx, fn, n, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / n * 1e3


def c3(ctx, frames):
    W, H, N = 7680, 4320, 100000
    objs, n, nbg, e, p = scene.random_scene(W, H, N).arrays()
    ctx.fb_configure(W, H)
    t = time.perf_counter()
    sc = ctx.scene_create(objs, nbg, e, p)
    ctx.sync()
    create_ms = (time.perf_counter() - t) * 1e3   # the first creation also grows the device memory pool
    again = []
    for _ in range(3):
        ctx.scene_free(sc)
        t = time.perf_counter()
        sc = ctx.scene_create(objs, nbg, e, p)
        ctx.sync()
        again.append((time.perf_counter() - t) * 1e3)
    ms = timed_frames(ctx, lambda: ctx.render_frame(sc, (0, 0, W, H)), frames)
    ctx.set_timing(True)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    walk, binning, _ = ctx.get_timing()
    ctx.set_timing(False)
    img = ctx.fb_read_rgba(0, 0, W, H)
    ctx.scene_free(sc)
    return {"config": "C3", "workload": f"{N} random layered polygons / brush strokes, {W}x{H}, seed 0xC0FFEE", "ms_per_frame": ms, "Mpx_per_s": W * H / ms / 1e3,
            "walker_ms": walk, "binning_ms": binning, "scene_create_ms": create_ms, "scene_recreate_ms": sorted(again)[1], "edges": int(len(e)), "stamp_points": int(len(p)),
            "checksum": int(img[::7, ::5].astype("uint64").sum())}


def c4(ctx, frames):
    W, H = 3840, 2160
    b, mover = scene.drag_scene(W, H, 7.0 * 0.45)
    objs, n, nbg, e, p = b.arrays()
    ctx.cache_clear()
    ctx.cache_configure(True, 100 << 20)  # engine.ml:1610-1611: 100 MiB
    ctx.fb_configure(W, H)
    sc = ctx.scene_create(objs, nbg, e, p)
    master = ctx.shape_box(0, 0, W, H)
    t = time.perf_counter()
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    first_ms = (time.perf_counter() - t) * 1e3
    px = [0]
    dirty_px = []

    def step():
        f = px[0]
        px[0] += 1
        dx, dy = round(3 * math.cos(2 * math.pi * f / 250)), round(2 * math.sin(2 * math.pi * f / 250))
        so, mo = ctx.scene_object_shape(sc, mover)
        ctx.scene_translate_object(sc, mover, dx, dy)
        sn, mn = ctx.scene_object_shape(sc, mover)
        dirty = ctx.dirty_region(so, mo, sn, mn, master, plain=False)  # a Group is Fancy: alldirty (render.ml:1396-1400)
        if len(dirty_px) < 50:
            dirty_px.append(ctx.shape_card(dirty))
        ctx.render_frame_shape(sc, dirty)
        for h in (so, mo, sn, mn, dirty):
            ctx.shape_free(h)

    ms = timed_frames(ctx, step, frames, warm=5)
    st = ctx.cache_stats()

    def step_fused():  # the same step through coh_scene_drag_object: no host round trip inside the frame
        f = px[0]
        px[0] += 1
        ctx.scene_drag_object(sc, mover, round(3 * math.cos(2 * math.pi * f / 250)), round(2 * math.sin(2 * math.pi * f / 250)))

    fused_ms = timed_frames(ctx, step_fused, frames, warm=5)
    sprite = ctx.cache_sprite_stats(sc)
    # the incrementally maintained framebuffer equals a fresh full render of the final scene
    inc = ctx.fb_read_rgba(0, 0, W, H)
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    same = bool(np.array_equal(inc, ctx.fb_read_rgba(0, 0, W, H)))
    full_ms = timed_frames(ctx, lambda: ctx.render_frame(sc, (0, 0, W, H)), 20)
    ctx.shape_free(master)
    ctx.scene_free(sc)
    return {"config": "C4", "workload": f"{frames}-frame drag of the lion group (alias path) over 400 static polygons, {W}x{H}, dirty region = alldirty from cached span sets",
            "ms_per_frame": fused_ms, "ms_per_frame_stepwise_api": ms, "first_frame_ms": first_ms, "full_frame_ms": full_ms, "mean_dirty_pixels": float(np.mean(dirty_px)), "cache": st, "sprite_cache": sprite,
            "incremental_equals_full_render": same}


def c5(ctx, frames):
    W, H = 3840, 2160
    objs, n, nbg, e, p = scene.filter_scene(W, H, 7.0).arrays()
    ctx.fb_configure(W, H)
    t = time.perf_counter()
    sc = ctx.scene_create(objs, nbg, e, p)
    ctx.sync()
    create_ms = (time.perf_counter() - t) * 1e3
    l0 = ctx.launch_count()
    ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync()
    launches = ctx.launch_count() - l0
    ms = timed_frames(ctx, lambda: ctx.render_frame(sc, (0, 0, W, H)), frames)
    img = ctx.fb_read_rgba(0, 0, W, H)
    ctx.scene_free(sc)
    return {"config": "C5", "workload": f"lion under blur (gaussian 5) / monochrome / affine lenses + Convolved (gaussian 4) page shadow, {W}x{H}",
            "ms_per_frame": ms, "Mpx_per_s": W * H / ms / 1e3, "scene_create_ms": create_ms, "launches_per_frame": int(launches),
            "checksum": int(img[::7, ::5].astype("uint64").sum())}


def main():
    which = [a for a in sys.argv[1:] if a in ("c3", "c4", "c5")] or ["c3", "c4", "c5"]
    frames = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else None
    ctx = abi.Context(0)
    for c in which:
        n = frames or {"c3": 10, "c4": 1000, "c5": 20}[c]
        print(json.dumps({"c3": c3, "c4": c4, "c5": c5}[c](ctx, n)), flush=True)


if __name__ == "__main__":
    main()
