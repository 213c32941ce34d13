"""Debug helper: where does a filter frame differ from the oracle?  (run on the GPU box)"""
import sys

import numpy as np

sys.path.insert(0, ".")
from coherence_renderer_b200 import abi, scene as S  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests.test_gpu_parity import _filter_scene, _finish, _render_both  # noqa: E402

ctx = abi.Context()
W, H = 200, 160
kind = sys.argv[1] if len(sys.argv) > 1 else "hole"
kw = {"kernel": ("gaussian", 3)} if kind == "blur" else {}
for matte in (None, S.Fill.plain(S.dissolve(S.rgba8(255, 255, 255), 170))):
    b, _ = _filter_scene(kind, W, H, matte=matte, **kw)
    got, ref, got_u, ref_u = _render_both(ctx, O, _finish(b, W, H), W, H)
    d = np.argwhere(got != ref)
    print("matte", "opaque" if matte is None else "translucent", "diff pixels", len(d))
    for y, x in d[:12]:
        print("  (x=%d,y=%d) got %08x ref %08x" % (x, y, got[y, x], ref[y, x]))
    if len(d):
        ys, xs = d[:, 0], d[:, 1]
        print("  bbox x %d..%d y %d..%d" % (xs.min(), xs.max(), ys.min(), ys.max()))
