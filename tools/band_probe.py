"""Dev tool: time one band of the C2 frame on one GPU (what one rank of an N-GPU run computes, without the
peer stores): python tools/band_probe.py N [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coherence_renderer_b200 import abi, scene, bands
W, H = 3840, 2160
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ks = [int(sys.argv[2])] if len(sys.argv) > 2 else list(range(N))
objs, n, nbg, e, p = scene.lion_scene(W, H, 7.0).arrays()
ctx = abi.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for k in ks:
    y0, y1 = bands.band_rows(H, N, k)
    ctx.fb_configure(W, H, y0, y1)
    sc = ctx.scene_create(objs, nbg, e, p)
    for _ in range(5): ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync(); ctx.set_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for a, b in ev:
        flush.zero_(); a.record(stream); ctx.render_frame(sc, (0, 0, W, H)); b.record(stream)
    torch.cuda.synchronize()
    walk, binning, _ = ctx.get_timing(); ctx.set_timing(False)
    print(f"band {k}/{N} rows {y0}-{y1}: frame {sum(a.elapsed_time(b) for a, b in ev) / len(ev):.4f} ms, walker {walk:.4f}, binning {binning:.4f}")
    ctx.scene_free(sc)
