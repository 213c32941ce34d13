"""Dev tool: per-phase cycle breakdown of the walker (library built with -DCOH_PHASE_PROFILE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["COH_LIB_PATH"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "coherence_renderer_b200", "libcoh_phase.so")
from coherence_renderer_b200 import abi, scene
W, H = 3840, 2160
objs, n, nbg, edges, points = scene.lion_scene(W, H, 7.0).arrays()
ctx = abi.Context(0); ctx.fb_configure(W, H); sc = ctx.scene_create(objs, nbg, edges, points)
buf = (C.c_ulonglong * 16)()
for _ in range(3): ctx.render_frame(sc, (0, 0, W, H))
abi.lib().coh_phase_cycles(ctx._h, buf, 1)
N = 10
for _ in range(N): ctx.render_frame(sc, (0, 0, W, H))
abi.lib().coh_phase_cycles(ctx._h, buf, 1)
names = ["other/loop", "scan", "group transitions", "row setup (shfl)", "AA", "composite"]
tot = sum(buf[:6])
for nm, v in zip(names, buf[:6]): print(f"{nm:22s} {v/N/1e6:9.2f} Mcycles/frame {100*v/tot:5.1f}%")
print("total warp-cycles/frame (M):", tot / N / 1e6)
print("  AA scaled scan %.1f M, AA pixel loop %.1f M" % (buf[6]/N/1e6, buf[7]/N/1e6))
print("  AA calls/frame %.0f, candidates/call %.2f, edge px/call %.2f, cycles/call scan %.0f, fallbacks/frame %.0f" % (buf[8]/N, buf[9]/max(buf[8],1), buf[10]/max(buf[8],1), buf[6]/max(buf[8],1), buf[11]/N))
print("  per call (cumulative cycles): staged-load %.0f, edge loop %.0f, finish %.0f" % (buf[12]/max(buf[8],1), buf[13]/max(buf[8],1), buf[14]/max(buf[8],1)))
import numpy as np
ncell = 120 * 135
cc = np.zeros(ncell, dtype=np.uint32)
abi.lib().coh_cell_cycles(ctx._h, cc.ctypes.data_as(C.POINTER(C.c_uint)), ncell)
cs = np.sort(cc)[::-1]
print("per-cell cycles: max %d, top10 %s, p99 %d, median %d, sum %.1f M" % (cs[0], cs[:10].tolist(), cs[ncell // 100], cs[ncell // 2], cc.sum() / 1e6))
