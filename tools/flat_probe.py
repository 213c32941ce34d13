"""Dev tool: C2 walker time with the lion inside its Group vs as flat top-level objects (cost of group transitions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
ctx = abi.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ctx.fb_configure(W, H)
for flat in (False, True):
    b = scene.lion_scene(W, H, 7.0)
    if flat:
        b.objs = [o for o in b.objs if o.kind not in (abi.COH_OBJ_GROUP_BEGIN, abi.COH_OBJ_GROUP_END)]
    objs, n, nbg, e, p = b.arrays()
    sc = ctx.scene_create(objs, nbg, e, p)
    for _ in range(5): ctx.render_frame(sc, (0, 0, W, H))
    ctx.sync(); ctx.set_timing(True)
    for _ in range(30):
        flush.zero_(); ctx.render_frame(sc, (0, 0, W, H))
    torch.cuda.synchronize()
    print("flat" if flat else "grouped", ctx.get_timing()); ctx.set_timing(False)
    ctx.scene_free(sc)
