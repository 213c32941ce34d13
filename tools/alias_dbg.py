import sys, numpy as np
sys.path.insert(0,'/root/repo')
from coherence_renderer_b200 import abi, scene as S
from oracle import pyoracle as O
W,H=400,300
pts = [(60.3, 40.2), (180.9, 70.1), (150.0, 200.7), (40.0, 160.0)]
ctx=abi.Context(0)
for (pd,bd,order) in (((-45,-30),(-45,-30),"pb"),((0,64),(0,64),"pb"),((-45,-30),(0,0),"pb"),((0,0),(-45,-30),"pb"),((0,0),(0,-30),"pb"),((0,0),(-45,0),"pb")):
    b=S.SceneBuilder()
    def poly(): b.polygon(pts, S.Fill.plain(S.dissolve(S.rgba8(30, 90, 200), 200)), dx=pd[0], dy=pd[1])
    def brush(): b.brush(0.9, 5.0, [[("C", (20.0, 250.0), (120.0, 20.0), (260.0, 280.0), (380.0, 60.0))]], S.Fill.plain(S.rgba8(200, 40, 40)), dx=bd[0], dy=bd[1])
    (poly(),brush()) if order=="pb" else (brush(),poly())
    b.begin_background(); b.rectangle(S.LIGHTGREY,0.,0.,float(W),float(H))
    objs,n,nbg,e,p=b.arrays()
    ref=O.render_frame(objs,n-nbg,nbg,e,p,(0,0,W,H))
    ctx.fb_configure(W,H); sc=ctx.scene_create(objs,nbg,e,p); ctx.render_frame(sc,(0,0,W,H)); ctx.sync()
    got=ctx.fb_read_rgba(0,0,W,H)
    d=np.abs(got.view(np.uint8).astype(int)-ref.view(np.uint8).astype(int)).reshape(H,W,4).max(axis=2)
    ys,xs=np.nonzero(d)
    print(pd,bd,order,"maxdiff",d.max(),"npix",len(ys), list(zip(ys[:6].tolist(),xs[:6].tolist())) if len(ys) else "")
    if len(ys):
        y,x=ys[0],xs[0]; print("   got",hex(got[y,x]),"ref",hex(ref[y,x]))
    ctx.scene_free(sc)
