"""Dev tool: one C2 frame as K bands on K streams of ONE device (K contexts sharing a framebuffer) against the whole
frame on one stream: do the latency-bound stages of one band hide behind the issue-bound ones of another?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
objs, n, nbg, edges, points = scene.lion_scene(W, H, 7.0).arrays()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
main = torch.cuda.current_stream()

def measure(cuts, steps=100, warm=10):
    K = len(cuts) - 1
    ctxs = [abi.Context(0) for _ in range(K)]
    for k, c in enumerate(ctxs):
        c.fb_configure(W, H, cuts[k], cuts[k + 1])
    for c in ctxs[1:]:
        c.fb_attach(ctxs[0].fb_device_ptr())
    scs = [c.scene_create(objs, nbg, edges, points) for c in ctxs]
    streams = [torch.cuda.ExternalStream(c.stream()) for c in ctxs]
    ms = []
    for s in range(warm + steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for k, c in enumerate(ctxs):
            streams[k].wait_event(e0)
            c.render_frame(scs[k], (0, cuts[k], W, cuts[k + 1] - cuts[k]))
            d = torch.cuda.Event(); d.record(streams[k]); main.wait_event(d)
        e1.record(main)
        torch.cuda.synchronize()
        if s >= warm:
            ms.append(e0.elapsed_time(e1))
    img = ctxs[0].fb_read_rgba(0, 0, W, H)
    chk = int(np.asarray(img).view(np.uint32).astype(np.uint64).sum())
    for c, sc in zip(ctxs, scs):
        c.scene_free(sc)
    for c in reversed(ctxs):
        c.close()
    ms.sort()
    return sum(ms) / len(ms), ms[len(ms) // 2], chk

for cuts in ([0, H], [0, 1088, H], [0, 720, 1440, H], [0, 544, 1088, 1632, H]):
    mean, med, chk = measure(cuts)
    print(f"bands {len(cuts) - 1}: mean {mean:.4f} ms  median {med:.4f} ms  checksum {chk}")
