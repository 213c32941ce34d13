"""Bands of the C3 frame on ONE GPU (what a rank of an N-GPU run renders, without peers): raster / binning time per
band, for the walker's work-item heights.  usage: python tools/c3_band_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, bands, scene

ctx = abi.Context(0)
W, H, N_OBJ = 7680, 4320, 100000
objs, n, nbg, e, p = scene.random_scene(W, H, N_OBJ).arrays()
ctx.fb_configure(W, H)
sc = ctx.scene_create(objs, nbg, e, p)
for wh in (0, 4, 1):
    ctx.set_option("walk_h", wh)
    for N in (1, 2, 4, 8):
        row = {"walk_h": wh, "bands": N, "raster_ms": [], "binning_ms": []}
        for k in sorted({0, N // 2, N - 1}):
            y0, y1 = bands.band_rows(H, N, k)
            ctx.fb_configure(W, H, y0, y1)
            for _ in range(3): ctx.render_frame(sc, (0, 0, W, H))
            ctx.set_timing(True)
            for _ in range(10): ctx.render_frame(sc, (0, 0, W, H))
            ctx.sync()
            walk, binning, _ = ctx.get_timing()
            ctx.set_timing(False)
            row["raster_ms"].append(round(walk, 4)); row["binning_ms"].append(round(binning, 4))
        print(json.dumps(row), flush=True)
