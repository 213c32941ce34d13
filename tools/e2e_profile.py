"""Dev tool: where the end-to-end (host buffers in, host pixels out) time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
objs, n, nbg, edges, points = scene.lion_scene(W, H, 7.0).arrays()
ctx = abi.Context(0); ctx.fb_configure(W, H)
host = np.zeros((H, W), dtype=np.uint32)
for it in range(4):
    t0 = time.perf_counter(); sc = ctx.scene_create(objs, nbg, edges, points); t1 = time.perf_counter()
    ctx.render_frame(sc, (0, 0, W, H)); ctx.sync(); t2 = time.perf_counter()
    ctx.fb_read_rgba(0, 0, W, H, out=host); t3 = time.perf_counter()
    ctx.scene_free(sc); t4 = time.perf_counter()
    print("scene_create %.2f ms, render+sync %.2f ms, read %.2f ms, free %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
