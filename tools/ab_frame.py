import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from coherence_renderer_b200 import abi, scene as S
ctx = abi.Context(0); W, H = 3840, 2160
objs, n, nbg, e, p = S.lion_scene(W, H, 7.0).arrays(); ctx.fb_configure(W, H); sc = ctx.scene_create(objs, nbg, e, p)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label):
    for _ in range(10): ctx.render_frame(sc, (0, 0, W, H))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for a, b in ev:
        flush.zero_(); a.record(stream); ctx.render_frame(sc, (0, 0, W, H)); b.record(stream)
    torch.cuda.synchronize()
    print(label, "cold %.4f" % (sum(a.elapsed_time(b) for a, b in ev) / 200), end=" ")
    ctx.sync(); t = time.perf_counter()
    for _ in range(500): ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync(); print("warm %.4f" % ((time.perf_counter() - t) / 500 * 1e3))
for rep in range(2):
    for ab in [int(x) for x in sys.argv[1:]] or [0]:
        try: ctx.set_option("ab", ab)
        except Exception: pass
        run(os.environ.get("COH_LIB_PATH", "new")[-10:] + " ab=%d" % ab)
