"""Phase times of coh_scene_create on the C3 scene (option "ab" bit 1 = trace on stderr)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coherence_renderer_b200 import abi, scene
ctx = abi.Context(0); W, H, N = 7680, 4320, 100000
objs, n, nbg, e, p = scene.random_scene(W, H, N).arrays()
ctx.fb_configure(W, H); ctx.set_option("ab", 2)
for i in range(3):
    t = time.perf_counter(); sc = ctx.scene_create(objs, nbg, e, p); ctx.sync(); print("scene_create total ms", (time.perf_counter() - t) * 1e3, flush=True); 
    if i < 2: ctx.scene_free(sc)
ctx.set_option("ab", 0)
for i in range(3):
    t = time.perf_counter(); s2 = ctx.scene_create(objs, nbg, e, p); ctx.sync(); print("untraced ms", (time.perf_counter() - t) * 1e3, flush=True); ctx.scene_free(s2)
ctx.render_frame(sc, (0, 0, W, H)); ctx.sync()
img = ctx.fb_read_rgba(0, 0, W, H); print("checksum", int(img[::7, ::5].astype("uint64").sum()))
