"""Dev tool: where the end-to-end frame (0.66 ms) exceeds the raw 33 MB read-back (0.585 ms)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
objs, n, nbg, edges, points = scene.lion_scene(W, H, 7.0).arrays()
ctx = abi.Context(0); ctx.fb_configure(W, H)
hosts = [torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
hn = [h.numpy().view(np.uint32) for h in hosts]
sc0 = ctx.scene_create(objs, nbg, edges, points)
ctx.render_frame(sc0, (0, 0, W, H)); ctx.sync()

def run(tag, body, k):
    for i in range(3): body(i)
    ctx.fb_read_wait(); ctx.sync()
    t = time.perf_counter()
    for i in range(k): body(i)
    ctx.fb_read_wait()
    s = (time.perf_counter() - t) / k
    print(f"{tag:60s} {k:4d} frames  {s*1e3:.4f} ms/frame  ({W*H*4/s/1e9:.1f} GB/s)")

def a(i): ctx.fb_read_rgba_async(0, 0, W, H, hn[i & 1])
def b(i): ctx.render_frame(sc0, (0, 0, W, H)); ctx.fb_read_rgba_async(0, 0, W, H, hn[i & 1])
def c(i):
    sh = ctx.scene_create(objs, nbg, edges, points); ctx.render_frame(sh, (0, 0, W, H)); ctx.fb_read_rgba_async(0, 0, W, H, hn[i & 1]); ctx.scene_free(sh)
for k in (20, 100):
    run("read-back only (staged, async)", a, k)
    run("render + read-back (scene resident)", b, k)
    run("scene_create + render + read-back + scene_free (bench e2e)", c, k)
# the pieces on the host clock
t = time.perf_counter()
for i in range(50):
    sh = ctx.scene_create(objs, nbg, edges, points); ctx.scene_free(sh)
ctx.sync()
print("scene_create + free alone: %.4f ms" % ((time.perf_counter() - t) / 50 * 1e3))
